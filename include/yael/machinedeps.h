/* yael/machinedeps.h -- drop-in prototypes (replaces /root/reference/yael/machinedeps.h:21-58).
 * The CPU thread pool has no role on the GPU path; compute_tasks is kept for callers. */
#ifndef YAEL_B200_MACHINEDEPS_H
#define YAEL_B200_MACHINEDEPS_H
#ifdef __cplusplus
extern "C" {
#endif
int count_cpu(void);        /* machinedeps.c:14-43: YAEL_COUNT_CPU or the affinity mask */
double getmillisecs(void);  /* machinedeps.c:91-96 */
void compute_tasks(int n, int nthread, void (*task_fun)(void *arg, int tid, int i),
                   void *task_arg); /* machinedeps.c:121-131 */
#ifdef __cplusplus
}
#endif
#endif
