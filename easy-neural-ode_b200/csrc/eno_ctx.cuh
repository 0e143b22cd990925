// The library context behind the opaque eno_ctx of include/eno.h, shared by the translation units of libeno.so.
#pragma once
#include <string>

#include "dist.cuh"
#include "eno_common.cuh"

struct eno_ctx {
  int device = 0;
  eno::Ctrl* mailbox = nullptr;  // pinned host copy of the device controller
  std::string err;
  int hint_fwd = 8;         // launches to enqueue before the first poll
  int hint_bwd = 8;
  int sm_count = 148;
  int path_mode = 0;        // 0 auto, 1 force streamed weights, 2 force cluster-resident (ENO_PATH)
  eno::Profiler prof;
  eno::DistCtx dist;             // world 1 unless eno_comm_init was called
  eno::Tableau* tabs = nullptr;  // pinned copies of all tableaux (sources of the constant-memory upload)
  int deferred = 0;         // 1: enqueue a fixed number of iterations, never synchronise (graph capturable)
  int spec_fwd = 0, spec_bwd = 0;
  void* css = nullptr;      // state of the GEMM-pipeline engine (css_engine.cu), created on first use
};
