/* integration/include/bwa/bwamem.h -- the three names bioseqdb's unchanged sources take from libbwa's bwamem.h / bntseq.h:
 *   bntamb1_t      the 16-byte hole record inside a NUCLSEQ datum      (reference bioseqdb/sequence.h:13-14,23-27)
 *   nst_nt4_table  letter -> code table used by nuclcode_from_char     (reference bioseqdb/sequence.h:51-53)
 *   mem_opt_t      the option struct extension.cpp writes by field     (reference bioseqdb/extension.cpp:220-231)
 * mem_opt_t keeps libbwa's field names for the 12 options the SQL composite carries; the drop-in bwa.cpp copies them into the
 * bsq_opts of the C ABI when the index is built / a read is aligned. */
#ifndef BIOSEQDB_GPU_BWA_BWAMEM_H
#define BIOSEQDB_GPU_BWA_BWAMEM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { int64_t offset; int32_t len; char amb; } bntamb1_t;
extern unsigned char nst_nt4_table[256];
typedef struct {
    int a, b;
    int o_del, e_del;
    int o_ins, e_ins;
    int pen_clip5, pen_clip3;
    int w;
    int zdrop;
    int min_seed_len;
    int max_occ;
} mem_opt_t;
#ifdef __cplusplus
}
#endif
#endif
