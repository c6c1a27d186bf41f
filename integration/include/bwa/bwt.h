/* integration/include/bwa/bwt.h -- what bioseqdb's own sources use from libbwa's bwt.h once the FM-index lives on the GPU: the two
 * integer typedefs (reference bioseqdb/sequence.h:7, bwa.h:9).  Nothing else of libbwa is needed: bwa.cpp no longer calls it. */
#ifndef BIOSEQDB_GPU_BWA_BWT_H
#define BIOSEQDB_GPU_BWA_BWT_H
#include <stdint.h>
typedef unsigned char ubyte_t;
typedef uint64_t bwtint_t;
#endif
