/* rb_io.c — saved nets: the reference's CDB "save format version 10".
 *
 * File layout written by reference recur-nn-io.c:12-147 and read back by
 * recur-nn-io.c:149-357: a constant database (rb_cdb.h) whose values are the
 * raw host-endian C objects, keyed "net.<field>", "bptt.<field>" and
 * "bottom_layer.<field>".  Keys and their order are kept so that files made
 * here and files made by the reference are interchangeable (tests/test_host_logic.py
 * checks both directions and the reference's own fixture
 * test/multi-text-6c34c563i73-h99-o3650.net).
 *
 * Weights are stored in the padded in-memory layout (ih_size / ho_size
 * floats).  Momentum, deltas, history and hidden state are not saved (format
 * version 6 and later).
 */
#include "rb_internal.h"
#include "rb_host.h"
#include "rb_cdb.h"
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define FORMAT_KEY "save_format_version"
#define MAX_METADATA_BYTES (100u * 1000u * 1000u)

static int
put(rb_cdb_writer *w, const char *key, const void *data, size_t bytes)
{
  int r = rb_cdb_writer_put(w, key, (uint32_t)strlen(key), data, (uint32_t)bytes);
  if (r)
    fprintf(stderr, "error %d saving '%s'\n", r, key);
  return r;
}

#define PUT_FIELD(obj, prefix, field) \
  put(&w, prefix "." #field, &(obj)->field, sizeof((obj)->field))

int
rnn_save_net(RecurNN *net, const char *filename, int backup)
{
  rb_cdb_writer w;
  char tmpfn[] = "tmp_net_XXXXXX";
  int fd = -1;
  if (net == NULL || filename == NULL)
    goto early_error;
  rb_host_will_touch_matrices(net);
  fd = mkostemp(tmpfn, O_RDWR | O_CREAT);
  if (fd == -1) {
    perror("can't open temporary file for writing");
    goto early_error;
  }
  if (rb_cdb_writer_begin(&w, fd))
    goto error;
  {
    const int version = 10;
    int bad = put(&w, FORMAT_KEY, &version, sizeof(version));
    bad |= PUT_FIELD(net, "net", i_size);
    bad |= PUT_FIELD(net, "net", h_size);
    bad |= PUT_FIELD(net, "net", o_size);
    bad |= PUT_FIELD(net, "net", input_size);
    bad |= PUT_FIELD(net, "net", hidden_size);
    bad |= PUT_FIELD(net, "net", output_size);
    bad |= PUT_FIELD(net, "net", ih_size);
    bad |= PUT_FIELD(net, "net", ho_size);
    bad |= PUT_FIELD(net, "net", generation);
    bad |= PUT_FIELD(net, "net", flags);
    bad |= PUT_FIELD(net, "net", presynaptic_noise);
    bad |= PUT_FIELD(net, "net", activation);
    bad |= PUT_FIELD(net, "net", rng);
    if (bad)
      goto error;
    if (put(&w, "net.ih_weights", net->ih_weights, sizeof(float) * net->ih_size) ||
        put(&w, "net.ho_weights", net->ho_weights, sizeof(float) * net->ho_size))
      goto error;
    if (net->metadata &&
        put(&w, "net.metadata", net->metadata, strlen(net->metadata) + 1))
      goto error;
    if ((net->flags & RNN_NET_FLAG_OWN_BPTT) && net->bptt) {
      RecurNNBPTT *bptt = net->bptt;
      bad |= PUT_FIELD(bptt, "bptt", depth);
      bad |= PUT_FIELD(bptt, "bptt", index);
      bad |= PUT_FIELD(bptt, "bptt", learn_rate);
      bad |= PUT_FIELD(bptt, "bptt", ho_scale);
      bad |= PUT_FIELD(bptt, "bptt", momentum);
      bad |= PUT_FIELD(bptt, "bptt", momentum_weight);
      bad |= PUT_FIELD(bptt, "bptt", min_error_factor);
      if (bad)
        goto error;
    }
    if (net->bottom_layer) {
      RecurExtraLayer *bottom_layer = net->bottom_layer;
      size_t matrix = (size_t)bottom_layer->i_size * bottom_layer->o_size;
      bad |= PUT_FIELD(bottom_layer, "bottom_layer", input_size);
      bad |= PUT_FIELD(bottom_layer, "bottom_layer", output_size);
      bad |= PUT_FIELD(bottom_layer, "bottom_layer", i_size);
      bad |= PUT_FIELD(bottom_layer, "bottom_layer", o_size);
      bad |= PUT_FIELD(bottom_layer, "bottom_layer", learn_rate_scale);
      bad |= PUT_FIELD(bottom_layer, "bottom_layer", overlap);
      bad |= put(&w, "bottom_layer.weights", bottom_layer->weights, sizeof(float) * matrix);
      if (bad)
        goto error;
    }
  }
  if (rb_cdb_writer_commit(&w)) {
    close(fd);
    unlink(tmpfn);
    goto early_error;
  }
  close(fd);
  if (backup) {
    /* Kept as the reference has it (recur-nn-io.c:126-135): the backup
       rename only happens when asprintf's length equals
       strlen(filename + 2), which never holds, so no "name~" file appears. */
    char *backup_filename;
    int size = asprintf(&backup_filename, "%s~", filename);
    if (size != -1) {
      if (size == (int)strlen(filename + 2))
        rename(filename, backup_filename);
      free(backup_filename);
    }
  }
  rename(tmpfn, filename);
  return 0;
error:
  rb_cdb_writer_abandon(&w);
  close(fd);
  unlink(tmpfn);
early_error:
  fprintf(stderr, "failed to save net %p with fd %d errno %d filename '%s'\n",
      (void *)net, fd, errno, filename ? filename : "(nil, which is the problem)");
  return -1;
}

/* Read one fixed-size value; 1 ok, 0 missing or wrong size. */
static int
get(int fd, const char *key, void *dest, size_t bytes)
{
  uint32_t vlen = 0;
  int r = rb_cdb_find(fd, key, (uint32_t)strlen(key), &vlen);
  if (r < 1) {
    fprintf(stderr, "error %d loading '%s'\n", r, key);
    return 0;
  }
  if (vlen != bytes) {
    fprintf(stderr, "size mismatch on '%s' want %zu, found %u\n", key, bytes, vlen);
    return 0;
  }
  return rb_cdb_read(fd, dest, vlen) == 0;
}

#define GET_FIELD(obj, prefix, bare, field) \
  get(fd, (version >= 4) ? prefix "." bare : bare, &(obj)->field, sizeof((obj)->field))

RecurNN *
rnn_load_net(const char *filename)
{
  RecurNN t;        /* scalar fields as saved */
  RecurNNBPTT tb;
  RecurExtraLayer tl;
  RecurNN *net = NULL;
  memset(&t, 0, sizeof(t));
  memset(&tb, 0, sizeof(tb));
  memset(&tl, 0, sizeof(tl));

  int fd = open(filename, O_RDONLY);
  if (fd == -1) {
    fprintf(stderr, "can't open '%s' (%s)\n", filename, strerror(errno));
    goto open_error;
  }
  int version = 0;
  {
    uint32_t vlen = 0;
    if (rb_cdb_find(fd, FORMAT_KEY, (uint32_t)strlen(FORMAT_KEY), &vlen) > 0 &&
        vlen == sizeof(version))
      rb_cdb_read(fd, &version, vlen);
  }
  if (!(GET_FIELD(&t, "net", "i_size", i_size) &&
        GET_FIELD(&t, "net", "h_size", h_size) &&
        GET_FIELD(&t, "net", "o_size", o_size) &&
        GET_FIELD(&t, "net", "input_size", input_size) &&
        GET_FIELD(&t, "net", "hidden_size", hidden_size) &&
        GET_FIELD(&t, "net", "output_size", output_size) &&
        GET_FIELD(&t, "net", "ih_size", ih_size) &&
        GET_FIELD(&t, "net", "ho_size", ho_size) &&
        GET_FIELD(&t, "net", "rng", rng) &&
        GET_FIELD(&t, "net", "generation", generation) &&
        GET_FIELD(&t, "net", "flags", flags)))
    goto pre_alloc_error;
  if (version >= 9) {
    if (!GET_FIELD(&t, "net", "presynaptic_noise", presynaptic_noise))
      goto pre_alloc_error;
  }
  else
    t.presynaptic_noise = 0;
  if (version >= 10) {
    if (!GET_FIELD(&t, "net", "activation", activation))
      goto pre_alloc_error;
  }
  else
    t.activation = RNN_RELU;

  if (t.flags & RNN_NET_FLAG_OWN_BPTT) {
    if (!(GET_FIELD(&tb, "bptt", "depth", depth) &&
          GET_FIELD(&tb, "bptt", "learn_rate", learn_rate) &&
          GET_FIELD(&tb, "bptt", "index", index) &&
          GET_FIELD(&tb, "bptt", "momentum", momentum) &&
          GET_FIELD(&tb, "bptt", "momentum_weight", momentum_weight)))
      goto pre_alloc_error;
    if (version >= 2) {
      if (!GET_FIELD(&tb, "bptt", "ho_scale", ho_scale))
        goto pre_alloc_error;
    }
    else
      tb.ho_scale = ((float)t.output_size) / t.hidden_size;
    if (version >= 3) {
      if (!GET_FIELD(&tb, "bptt", "min_error_factor", min_error_factor))
        goto pre_alloc_error;
    }
    else
      tb.min_error_factor = BASE_MIN_ERROR_FACTOR * t.h_size;
  }
  if ((t.flags & RNN_NET_FLAG_BOTTOM_LAYER) && version >= 4) {
    if (!(GET_FIELD(&tl, "bottom_layer", "learn_rate_scale", learn_rate_scale) &&
          GET_FIELD(&tl, "bottom_layer", "input_size", input_size) &&
          GET_FIELD(&tl, "bottom_layer", "output_size", output_size) &&
          GET_FIELD(&tl, "bottom_layer", "i_size", i_size) &&
          GET_FIELD(&tl, "bottom_layer", "o_size", o_size) &&
          GET_FIELD(&tl, "bottom_layer", "overlap", overlap)))
      goto pre_alloc_error;
  }

  if (t.flags & RNN_NET_FLAG_BOTTOM_LAYER) {
    net = rnn_new_with_bottom_layer(tl.input_size, tl.output_size, t.hidden_size,
        t.output_size, t.flags, 0, NULL, tb.depth, tb.learn_rate, tb.momentum,
        t.presynaptic_noise, t.activation, tl.overlap);
  }
  else {
    net = rnn_new(t.input_size, t.hidden_size, t.output_size, t.flags, 0, NULL,
        tb.depth, tb.learn_rate, tb.momentum, t.presynaptic_noise, t.activation);
  }
  net->rng = t.rng;
  net->generation = t.generation;
  if (net->bptt) {
    /* As in the reference (recur-nn-io.c:249-254) the saved ring index is
       installed without re-pointing input_layer/real_inputs; the next
       rnn_bptt_advance re-derives both from it. */
    net->bptt->index = tb.index;
    net->bptt->momentum_weight = tb.momentum_weight;
    net->bptt->ho_scale = tb.ho_scale;
    net->bptt->min_error_factor = tb.min_error_factor;
  }
  /* the saved sizes must be what rnn_new derives from the three basic ones */
  if (net->i_size != t.i_size || net->h_size != t.h_size || net->o_size != t.o_size ||
      net->ih_size != t.ih_size || net->ho_size != t.ho_size ||
      net->activation != t.activation) {
    fprintf(stderr, "saved sizes of '%s' do not match a freshly made net\n", filename);
    goto error;
  }
  {
    uint32_t vlen = 0;
    const char *key = (version >= 4) ? "net.ih_weights" : "ih_weights";
    if (rb_cdb_find(fd, key, (uint32_t)strlen(key), &vlen) < 1 ||
        vlen != net->ih_size * sizeof(float) ||
        rb_cdb_read(fd, net->ih_weights, vlen)) {
      fprintf(stderr, "cannot load '%s' (%u bytes)\n", key, vlen);
      goto error;
    }
    key = (version >= 4) ? "net.ho_weights" : "ho_weights";
    if (rb_cdb_find(fd, key, (uint32_t)strlen(key), &vlen) < 1 ||
        vlen != net->ho_size * sizeof(float) ||
        rb_cdb_read(fd, net->ho_weights, vlen)) {
      fprintf(stderr, "cannot load '%s' (%u bytes)\n", key, vlen);
      goto error;
    }
    if (version >= 5) {
      int r = rb_cdb_find(fd, "net.metadata", 12, &vlen);
      if (r < 1) {
        fprintf(stderr, "error %d loading 'net.metadata'\ncontinuing anyway\n", r);
      }
      else {
        if (vlen > MAX_METADATA_BYTES) {
          fprintf(stderr, "size of 'net.metadata'(%u) exceeds maximum %u\n", vlen,
              MAX_METADATA_BYTES);
          goto error;
        }
        net->metadata = (char *)malloc(vlen + 1);
        if (!net->metadata || rb_cdb_read(fd, net->metadata, vlen))
          goto error;
        net->metadata[vlen] = 0;
      }
    }
    if (net->bottom_layer) {
      RecurExtraLayer *bl = net->bottom_layer;
      size_t want = (size_t)bl->i_size * bl->o_size * sizeof(float);
      if (rb_cdb_find(fd, "bottom_layer.weights", 20, &vlen) < 1 || vlen != want ||
          rb_cdb_read(fd, bl->weights, vlen)) {
        fprintf(stderr, "cannot load 'bottom_layer.weights'\n");
        goto error;
      }
      /* like the reference (recur-nn-io.c:225,292-299: read, never
         installed), the saved learn_rate_scale does not override the
         constructor's default */
    }
  }
  close(fd);
  rb_weights_changed(net);
  fprintf(stderr, "successfully loaded net '%s'\n", filename);
  return net;
error:
  rnn_delete_net(net);
pre_alloc_error:
  close(fd);
open_error:
  fprintf(stderr, "loading net failed!\n");
  return NULL;
}
