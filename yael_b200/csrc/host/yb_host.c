/* yb_host.c -- staging helpers of the C host layer. */
#include "yb_host.h"

#include <stdio.h>
#include <stdlib.h>

void ybh_die(const char *where, int rc) {
  fprintf(stderr, "yael_b200: %s failed (code %d): %s\n", where, rc, yb_last_error());
  abort();
}

int ybh_is_device_ptr(const void *p) { return yb_is_device_ptr(p); }

ybh_arg ybh_in(const void *p, size_t bytes) {
  ybh_arg a = {0, 0, bytes, 0};
  if (!p || bytes == 0) return a;
  if (yb_is_device_ptr(p)) {
    a.dev = (void *)p;
    return a;
  }
  a.dev = yb_malloc(bytes);
  a.host = (void *)p;
  a.owned = 1;
  YBH_CHECK(yb_h2d(a.dev, p, bytes, NULL));
  return a;
}

ybh_arg ybh_out(void *p, size_t bytes) {
  ybh_arg a = {0, 0, bytes, 0};
  if (!p || bytes == 0) return a;
  if (yb_is_device_ptr(p)) {
    a.dev = p;
    return a;
  }
  a.dev = yb_malloc(bytes);
  a.host = p;
  a.owned = 1;
  return a;
}

void ybh_finish(ybh_arg *a, int copy_back) {
  if (a->owned) {
    if (copy_back && a->host) {
      YBH_CHECK(yb_d2h(a->host, a->dev, a->bytes, NULL));
      YBH_CHECK(yb_sync(NULL));
    }
    yb_free(a->dev);
  }
  a->dev = 0;
  a->owned = 0;
}

void ybh_sync(void) { YBH_CHECK(yb_sync(NULL)); }
