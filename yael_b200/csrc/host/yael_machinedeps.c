/* yael_machinedeps.c -- include/yael/machinedeps.h (yael/machinedeps.c:14-131). */
#define _GNU_SOURCE
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>

#include "../../../include/yael/machinedeps.h"

int count_cpu(void) { /* machinedeps.c:14-43 */
  const char *e = getenv("YAEL_COUNT_CPU");
  if (e) {
    int n;
    if (sscanf(e, "%d", &n) == 1 && n > 0) return n;
    fprintf(stderr, "could not parse YAEL_CPU_COUNT environment variable, using default\n");
  }
  cpu_set_t set;
  sched_getaffinity(0, sizeof(set), &set);
  return CPU_COUNT(&set);
}

double getmillisecs(void) { /* machinedeps.c:91-96 */
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

void compute_tasks(int n, int nt, void (*task_fun)(void *arg, int tid, int i), void *task_arg) {
  /* machinedeps.c:121-131: an OpenMP dynamic loop in the reference; the library itself no
   * longer needs host threads, so tasks run in order on the calling thread */
  (void)nt;
  for (int i = 0; i < n; i++) (*task_fun)(task_arg, 0, i);
}
