// vvcb_broker -- the server process of include/vvc_intra_b200_broker.h: one per GPU, owns the engine context that the walker
// processes (VVCB_BROKER=<path> in their environment) share.
//   vvcb_broker <path> [--device D] [--bit-depth B] [--ctu C] [--clients N] [--frame WxH] [--workers T]
// Runs until `vvcb_broker <path> --stop` (or vvcb_broker_stop from any process).  `--stats` prints the counters as JSON.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/vvc_intra_b200_broker.h"

int main(int argc, char** argv)
{
  if (argc < 2) { fprintf(stderr, "usage: vvcb_broker <path> [--device D] [--bit-depth B] [--ctu C] [--clients N] [--frame WxH] [--workers T] | --stop | --stats\n"); return 2; }
  const char* path = argv[1];
  int device = 0, bd = 10, ctu = 128, clients = 64, fw = 1920, fh = 1080, workers = 4;
  for (int i = 2; i < argc; i++) {
    if (!strcmp(argv[i], "--stop")) return vvcb_broker_stop(path) == VVCB_OK ? 0 : 1;
    if (!strcmp(argv[i], "--stats")) {
      vvcb_broker_stats s;
      if (vvcb_broker_read_stats(path, &s) != VVCB_OK) return 1;
      printf("{\"cycles\": %llu, \"requests\": %llu, \"cu_requests\": %llu, \"visits\": %llu, \"tu_jobs\": %llu, \"max_batch\": %llu, \"busy_ns\": %llu, \"wall_ns\": %llu, "
             "\"clients_seen\": %llu, \"kernel_launches\": %llu, \"phase_ns\": [%llu, %llu, %llu, %llu, %llu, %llu, %llu, %llu]}\n",
             (unsigned long long)s.cycles, (unsigned long long)s.requests, (unsigned long long)s.cu_requests, (unsigned long long)s.visits, (unsigned long long)s.tu_jobs,
             (unsigned long long)s.max_batch, (unsigned long long)s.busy_ns, (unsigned long long)s.wall_ns, (unsigned long long)s.clients_seen, (unsigned long long)s.kernel_launches,
             (unsigned long long)s.phase_ns[0], (unsigned long long)s.phase_ns[1], (unsigned long long)s.phase_ns[2], (unsigned long long)s.phase_ns[3], (unsigned long long)s.phase_ns[4], (unsigned long long)s.phase_ns[5], (unsigned long long)s.phase_ns[6], (unsigned long long)s.phase_ns[7]);
      return 0;
    }
    if (i + 1 >= argc) { fprintf(stderr, "vvcb_broker: %s needs a value\n", argv[i]); return 2; }
    if (!strcmp(argv[i], "--device")) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--bit-depth")) bd = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--ctu")) ctu = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--clients")) clients = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--workers")) workers = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--frame")) { if (sscanf(argv[++i], "%dx%d", &fw, &fh) != 2) { fprintf(stderr, "vvcb_broker: --frame WxH\n"); return 2; } }
    else { fprintf(stderr, "vvcb_broker: unknown option %s\n", argv[i]); return 2; }
  }
  const int rc = vvcb_broker_serve(path, device, bd, ctu, clients, fw, fh, workers);
  if (rc != VVCB_OK) fprintf(stderr, "vvcb_broker: serve failed (%d): %s\n", rc, vvcb_last_error(nullptr));
  return rc == VVCB_OK ? 0 : 1;
}
