/* oracle/shim/cdb.h — TEST INFRASTRUCTURE ONLY.
 *
 * The reference's recur-nn-io.c includes <cdb.h> from tinycdb, which is
 * neither vendored in /root/reference nor installed in this image.  This
 * header supplies the five tinycdb entry points recur-nn-io.c calls
 * (recur-nn-io.c:26,45-60,124 cdb_make_*; 168-184 cdb_seek/cdb_bread) on
 * top of this repo's own constant-database code, so that the UNMODIFIED
 * reference sources compile into oracle/_ref/ (see oracle/Makefile).
 */
#ifndef ORACLE_SHIM_CDB_H
#define ORACLE_SHIM_CDB_H

#include "rb_cdb.h"

struct cdb_make {
  rb_cdb_writer w;
};

static inline int
cdb_make_start(struct cdb_make *c, int fd)
{
  return rb_cdb_writer_begin(&c->w, fd);
}

static inline int
cdb_make_add(struct cdb_make *c, const void *key, unsigned klen,
    const void *val, unsigned vlen)
{
  return rb_cdb_writer_put(&c->w, key, klen, val, vlen);
}

static inline int
cdb_make_finish(struct cdb_make *c)
{
  return rb_cdb_writer_commit(&c->w);
}

static inline int
cdb_seek(int fd, const void *key, unsigned klen, unsigned *vlen)
{
  uint32_t v = 0;
  int r = rb_cdb_find(fd, key, klen, &v);
  if (r > 0 && vlen)
    *vlen = v;
  return r;
}

static inline int
cdb_bread(int fd, void *buf, int len)
{
  return rb_cdb_read(fd, buf, (uint32_t)len);
}

#endif
