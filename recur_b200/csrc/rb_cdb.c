/* rb_cdb.c — constant-database reader/writer (format notes in rb_cdb.h). */
#include "rb_cdb.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define RB_CDB_HEADER_BYTES 2048u

static inline void
put_le32(uint8_t *p, uint32_t v)
{
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
  p[2] = (uint8_t)(v >> 16);
  p[3] = (uint8_t)(v >> 24);
}

static inline uint32_t
get_le32(const uint8_t *p)
{
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
    ((uint32_t)p[3] << 24);
}

uint32_t
rb_cdb_hash(const void *key, uint32_t klen)
{
  const uint8_t *k = (const uint8_t *)key;
  uint32_t h = 5381;
  for (uint32_t i = 0; i < klen; i++) {
    h = (h * 33u) ^ k[i];
  }
  return h;
}

static int
write_fully(int fd, const void *buf, size_t len)
{
  const uint8_t *p = (const uint8_t *)buf;
  while (len) {
    ssize_t n = write(fd, p, len);
    if (n < 0) {
      if (errno == EINTR)
        continue;
      return -1;
    }
    p += n;
    len -= (size_t)n;
  }
  return 0;
}

static int
pread_fully(int fd, void *buf, size_t len, off_t at)
{
  uint8_t *p = (uint8_t *)buf;
  while (len) {
    ssize_t n = pread(fd, p, len, at);
    if (n < 0) {
      if (errno == EINTR)
        continue;
      return -1;
    }
    if (n == 0) {
      errno = EPROTO; /* truncated file */
      return -1;
    }
    p += n;
    at += n;
    len -= (size_t)n;
  }
  return 0;
}

int
rb_cdb_writer_begin(rb_cdb_writer *w, int fd)
{
  memset(w, 0, sizeof(*w));
  w->fd = fd;
  w->pos = RB_CDB_HEADER_BYTES;
  /* leave room for the header; it is written last */
  if (lseek(fd, RB_CDB_HEADER_BYTES, SEEK_SET) < 0) {
    w->failed = 1;
    return -1;
  }
  return 0;
}

int
rb_cdb_writer_put(rb_cdb_writer *w, const void *key, uint32_t klen,
    const void *val, uint32_t vlen)
{
  if (w->failed)
    return -1;
  uint64_t end = (uint64_t)w->pos + 8u + klen + vlen;
  if (end > 0xffffffffull) { /* the format is limited to 4 GiB */
    errno = EFBIG;
    w->failed = 1;
    return -1;
  }
  if (w->n_entries == w->cap_entries) {
    size_t cap = w->cap_entries ? w->cap_entries * 2 : 64;
    rb_cdb_entry *e = (rb_cdb_entry *)realloc(w->entries, cap * sizeof(*e));
    if (!e) {
      w->failed = 1;
      return -1;
    }
    w->entries = e;
    w->cap_entries = cap;
  }
  uint8_t head[8];
  put_le32(head, klen);
  put_le32(head + 4, vlen);
  if (write_fully(w->fd, head, 8) || write_fully(w->fd, key, klen) ||
      write_fully(w->fd, val, vlen)) {
    w->failed = 1;
    return -1;
  }
  w->entries[w->n_entries].hash = rb_cdb_hash(key, klen);
  w->entries[w->n_entries].pos = w->pos;
  w->n_entries++;
  w->pos = (uint32_t)end;
  return 0;
}

void
rb_cdb_writer_abandon(rb_cdb_writer *w)
{
  free(w->entries);
  w->entries = NULL;
  w->n_entries = w->cap_entries = 0;
  w->failed = 1;
}

int
rb_cdb_writer_commit(rb_cdb_writer *w)
{
  if (w->failed) {
    rb_cdb_writer_abandon(w);
    return -1;
  }
  uint32_t count[256];
  uint32_t start[256];
  uint8_t header[RB_CDB_HEADER_BYTES];
  memset(count, 0, sizeof(count));
  for (size_t i = 0; i < w->n_entries; i++) {
    count[w->entries[i].hash & 255u]++;
  }
  /* group the records by table, keeping insertion order inside a table */
  rb_cdb_entry *grouped =
    (rb_cdb_entry *)malloc((w->n_entries + 1) * sizeof(rb_cdb_entry));
  uint32_t biggest = 0;
  uint32_t run = 0;
  for (int t = 0; t < 256; t++) {
    start[t] = run;
    run += count[t];
    if (count[t] > biggest)
      biggest = count[t];
  }
  uint8_t *slots = (uint8_t *)malloc((size_t)biggest * 2u * 8u + 8u);
  if (!grouped || !slots) {
    free(grouped);
    free(slots);
    rb_cdb_writer_abandon(w);
    return -1;
  }
  {
    uint32_t fill[256];
    memcpy(fill, start, sizeof(fill));
    for (size_t i = 0; i < w->n_entries; i++) {
      grouped[fill[w->entries[i].hash & 255u]++] = w->entries[i];
    }
  }
  int ret = 0;
  for (int t = 0; t < 256 && ret == 0; t++) {
    uint32_t n_slots = count[t] * 2u;
    put_le32(header + t * 8, w->pos);
    put_le32(header + t * 8 + 4, n_slots);
    if (n_slots == 0)
      continue;
    memset(slots, 0, (size_t)n_slots * 8u);
    for (uint32_t j = 0; j < count[t]; j++) {
      const rb_cdb_entry *e = &grouped[start[t] + j];
      uint32_t s = (e->hash >> 8) % n_slots;
      while (get_le32(slots + s * 8 + 4) != 0) { /* record_pos 0 == empty */
        s = (s + 1 == n_slots) ? 0 : s + 1;
      }
      put_le32(slots + s * 8, e->hash);
      put_le32(slots + s * 8 + 4, e->pos);
    }
    uint64_t end = (uint64_t)w->pos + (uint64_t)n_slots * 8u;
    if (end > 0xffffffffull) {
      errno = EFBIG;
      ret = -1;
      break;
    }
    if (write_fully(w->fd, slots, (size_t)n_slots * 8u))
      ret = -1;
    w->pos = (uint32_t)end;
  }
  if (ret == 0) {
    if (lseek(w->fd, 0, SEEK_SET) < 0 ||
        write_fully(w->fd, header, sizeof(header)))
      ret = -1;
  }
  free(grouped);
  free(slots);
  free(w->entries);
  w->entries = NULL;
  w->n_entries = w->cap_entries = 0;
  if (ret)
    w->failed = 1;
  return ret;
}

int
rb_cdb_find(int fd, const void *key, uint32_t klen, uint32_t *vlen)
{
  uint8_t pair[8];
  uint32_t h = rb_cdb_hash(key, klen);
  if (pread_fully(fd, pair, 8, (off_t)((h & 255u) * 8u)))
    return -1;
  uint32_t table_pos = get_le32(pair);
  uint32_t n_slots = get_le32(pair + 4);
  if (n_slots == 0)
    return 0;
  uint32_t s = (h >> 8) % n_slots;
  for (uint32_t probes = 0; probes < n_slots; probes++) {
    if (pread_fully(fd, pair, 8, (off_t)table_pos + (off_t)s * 8))
      return -1;
    uint32_t rec_pos = get_le32(pair + 4);
    if (rec_pos == 0)
      return 0;
    if (get_le32(pair) == h) {
      uint8_t head[8];
      if (pread_fully(fd, head, 8, (off_t)rec_pos))
        return -1;
      uint32_t rk = get_le32(head);
      uint32_t rv = get_le32(head + 4);
      if (rk == klen) {
        int same = 1;
        uint8_t buf[256];
        uint32_t done = 0;
        while (done < klen && same) {
          uint32_t chunk = klen - done;
          if (chunk > sizeof(buf))
            chunk = sizeof(buf);
          if (pread_fully(fd, buf, chunk, (off_t)rec_pos + 8 + done))
            return -1;
          same = memcmp(buf, (const uint8_t *)key + done, chunk) == 0;
          done += chunk;
        }
        if (same) {
          if (lseek(fd, (off_t)rec_pos + 8 + klen, SEEK_SET) < 0)
            return -1;
          if (vlen)
            *vlen = rv;
          return 1;
        }
      }
    }
    s = (s + 1 == n_slots) ? 0 : s + 1;
  }
  return 0;
}

int
rb_cdb_read(int fd, void *buf, uint32_t len)
{
  uint8_t *p = (uint8_t *)buf;
  while (len) {
    ssize_t n = read(fd, p, len);
    if (n < 0) {
      if (errno == EINTR)
        continue;
      return -1;
    }
    if (n == 0) {
      errno = EPROTO;
      return -1;
    }
    p += n;
    len -= (uint32_t)n;
  }
  return 0;
}
