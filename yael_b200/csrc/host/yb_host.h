/* yb_host.h -- helpers shared by the C host layer (drop-in API on top of the yb_ C ABI). */
#ifndef YB_HOST_H
#define YB_HOST_H
#include <stddef.h>

#include "../../../include/yael_b200.h"

/* message on stderr + abort(): the reference's convention for unrecoverable conditions
 * (yael/vector.c:37-40); used when a yb_ call reports a CUDA failure */
void ybh_die(const char *where, int rc);
#define YBH_CHECK(call)                   \
  do {                                    \
    int _rc = (call);                     \
    if (_rc) ybh_die(#call, _rc);         \
  } while (0)

/* 1 when p is device (or managed) memory that kernels can read directly */
int ybh_is_device_ptr(const void *p);

/* An argument staged for the device: host pointers are copied into a pooled device buffer,
 * device pointers are used in place. */
typedef struct {
  void *dev;   /* pointer usable by kernels */
  void *host;  /* caller's pointer when it is host memory (NULL when dev is the caller's) */
  size_t bytes;
  int owned;   /* dev came from yb_malloc */
} ybh_arg;

ybh_arg ybh_in(const void *p, size_t bytes);  /* input: copies host -> device */
ybh_arg ybh_out(void *p, size_t bytes);       /* output: allocates, ybh_finish copies back */
void ybh_finish(ybh_arg *a, int copy_back);   /* D2H (if output) + release */
void ybh_sync(void);
#endif
