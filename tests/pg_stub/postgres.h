/* tests/pg_stub/postgres.h -- TEST INFRASTRUCTURE: the handful of PostgreSQL server symbols that reference bioseqdb/sequence.h,
 * sequence.cpp and the bwa part of extension.cpp touch, so that the drop-in adapter (integration/bioseqdb/bwa.{h,cpp}) can be
 * compiled and linked against the reference's UNCHANGED sources on a machine without PostgreSQL. */
#ifndef PG_STUB_POSTGRES_H
#define PG_STUB_POSTGRES_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef uintptr_t Datum;
typedef unsigned int Oid;
typedef size_t Size;
static inline void* palloc(Size n) { return malloc(n ? n : 1); }
static inline void* palloc0(Size n) { return calloc(n ? n : 1, 1); }
static inline void pfree(void* p) { free(p); }
#define SET_VARSIZE(ptr, len) (*(uint32_t*)(ptr) = ((uint32_t)(len)) << 2)
#define VARSIZE(ptr) ((*(const uint32_t*)(ptr)) >> 2)
#define VARHDRSZ 4
#endif
