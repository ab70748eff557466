/* rb_cdb — a minimal reader/writer for D. J. Bernstein's "constant database"
 * file format, the container recur uses for saved nets.
 *
 * The reference links tinycdb for this (recur-nn-io.c:3, Makefile:50-51);
 * tinycdb is not vendored in the reference and not installed here, so the
 * format is restated from its public description (cr.yp.to/cdb/cdb.txt):
 *
 *   [0, 2048)        256 x { u32 table_pos, u32 table_slots }   little endian
 *   records          { u32 klen, u32 dlen, key bytes, data bytes } ...
 *   256 hash tables  slots x { u32 hash, u32 record_pos }, open addressing,
 *                    table t holds the keys with (hash & 255) == t, a record
 *                    probes from slot (hash >> 8) % slots, and every table has
 *                    twice as many slots as records.
 *   hash             h = 5381; for each byte c: h = (h * 33) ^ c  (mod 2^32)
 *
 * The only on-disk fixture that pins this is the reference's
 * test/multi-text-6c34c563i73-h99-o3650.net (see tests/test_cdb.py).
 */
#ifndef RB_CDB_H
#define RB_CDB_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_cdb_entry {
  uint32_t hash;
  uint32_t pos;
} rb_cdb_entry;

typedef struct rb_cdb_writer {
  int fd;
  uint32_t pos;          /* next byte to be written */
  rb_cdb_entry *entries; /* one per record, in insertion order */
  size_t n_entries;
  size_t cap_entries;
  int failed;
} rb_cdb_writer;

uint32_t rb_cdb_hash(const void *key, uint32_t klen);

/* Writer: begin on an open, empty, seekable fd; put records; commit writes
   the hash tables and the 2048-byte header.  All return 0 on success, -1 on
   error (errno set where the OS reported one). */
int rb_cdb_writer_begin(rb_cdb_writer *w, int fd);
int rb_cdb_writer_put(rb_cdb_writer *w, const void *key, uint32_t klen,
    const void *val, uint32_t vlen);
int rb_cdb_writer_commit(rb_cdb_writer *w);
void rb_cdb_writer_abandon(rb_cdb_writer *w);

/* Reader: look the key up in the file behind fd.  Returns 1 and leaves the
   file offset at the first data byte (length in *vlen) when found, 0 when
   absent, -1 on I/O or format error. */
int rb_cdb_find(int fd, const void *key, uint32_t klen, uint32_t *vlen);

/* Read exactly len bytes from the current offset. 0 on success, -1 on error. */
int rb_cdb_read(int fd, void *buf, uint32_t len);

#ifdef __cplusplus
}
#endif
#endif
