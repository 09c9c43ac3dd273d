/*
 * vvc_intra_b200 broker -- several host walkers (encoder processes, one picture each) share one engine context per GPU.
 *
 * SURVEY.md 7.2 option (A): a bit-exact use of the engine is one small batch per estIntraPredLumaQT call
 * (EL/IntraSearch.cpp:289), far too small to fill a B200, and the reference is one single-threaded process per picture
 * (CL/TypeDef.h:318-330).  The broker is the piece in between: every walker process maps one shared-memory file, posts its
 * vvcb_cu_request there and sleeps; one server process per GPU owns the vvcb_ctx, collects whatever requests are pending,
 * runs them as ONE vvcb_cu_eval batch (the pictures of all clients live side by side in one plane) and wakes the walkers.
 *
 * Client side: nothing to call -- vvcb_create() returns a proxy context when the environment variable VVCB_BROKER names the
 * broker file; vvcb_frame_begin / vvcb_reco_update / vvcb_reco_update_rects / vvcb_rmd_eval / vvcb_cu_eval / vvcb_set_option
 * then travel through the broker, every other entry point fails with VVCB_ERR_STATE.
 */
#ifndef VVC_INTRA_B200_BROKER_H
#define VVC_INTRA_B200_BROKER_H
#include "vvc_intra_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vvcb_broker_stats {
  uint64_t cycles;            /* engine batches issued                                                        */
  uint64_t requests;          /* client round trips served                                                    */
  uint64_t cu_requests;       /* vvcb_cu_request entries inside them                                          */
  uint64_t visits, tu_jobs;   /* rough-mode-decision visits and TU jobs evaluated                             */
  uint64_t max_batch;         /* largest number of round trips merged into one batch                          */
  uint64_t busy_ns;           /* time the worker threads spent inside engine calls, summed over the workers   */
  uint64_t wall_ns;           /* wall time since the server started serving                                   */
  uint64_t clients_seen;      /* clients that ever connected                                                  */
  uint64_t kernel_launches;   /* vvcb_launch_count of the server's context                                    */
  uint64_t phase_ns[8];       /* vvcb_cu_eval_phases summed over the workers' contexts                        */
} vvcb_broker_stats;

/* Server: creates the broker file at `path`, `workers` engine contexts on `device` (one per worker thread, each with its own stream,
 * all sharing one plane that holds max_clients pictures of at most frame_width x frame_height), then serves until
 * vvcb_broker_stop(path) is called (from any process).  A worker claims the requests that are pending when it looks and runs them as
 * one batch; several workers keep several batches in flight, so a batch with a long serial chain (a 32x32 dependent quantisation)
 * does not hold up the walkers of the next one.  Blocking; returns VVCB_OK after a clean stop.                                     */
int vvcb_broker_serve(const char* path, int device, int bit_depth, int ctu_size, int max_clients, int frame_width, int frame_height, int workers);
int vvcb_broker_stop(const char* path);
int vvcb_broker_read_stats(const char* path, vvcb_broker_stats* out);   /* readable while serving and after the stop        */

#ifdef __cplusplus
}
#endif
#endif
