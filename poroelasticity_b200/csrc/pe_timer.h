// Phase timer (tracing aid): CUDA events between named marks on one stream, accumulated per name and
// printed per solve.  Enabled with PE_TIMING=1; only meaningful when the iteration is not replayed from a CUDA
// graph (multi-GPU default, or solver option bit 6).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges show up in nsys / ncu --nvtx timelines, no-ops otherwise

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace pe
{
  struct PhaseTimer
  {
    bool                                 on = false;
    std::vector<cudaEvent_t>             pool;
    std::vector<std::pair<int, std::string>> marks; // event index, name of the phase that ENDS at this mark
    size_t                               used = 0;
    std::map<std::string, double>        acc;
    std::map<std::string, int>           cnt;
    static PhaseTimer &
    get()
    {
      static PhaseTimer t;
      static bool       init = false;
      if (!init)
        {
          init = true;
          t.on = getenv("PE_TIMING") != nullptr;
        }
      return t;
    }
    void
    mark(cudaStream_t s, const char *name)
    {
      if (!on)
        return;
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &st);
      if (st != cudaStreamCaptureStatusNone)
        return;
      if (used == pool.size())
        {
          cudaEvent_t e;
          cudaEventCreate(&e);
          pool.push_back(e);
        }
      cudaEventRecord(pool[used], s);
      marks.emplace_back((int)used, name);
      ++used;
    }
    void
    flush(const int rank, const int iterations)
    {
      if (!on || marks.size() < 2)
        {
          marks.clear();
          used = 0;
          return;
        }
      cudaEventSynchronize(pool[marks.back().first]);
      for (size_t k = 1; k < marks.size(); ++k)
        {
          float ms = 0.f;
          cudaEventElapsedTime(&ms, pool[marks[k - 1].first], pool[marks[k].first]);
          acc[marks[k].second] += ms;
          cnt[marks[k].second] += 1;
        }
      const char *env_rank = getenv("RANK"); // one process per GPU under torchrun: only rank 0 prints
      if (rank == 0 && (!env_rank || atoi(env_rank) == 0))
        {
          double tot = 0.;
          for (auto &kv : acc)
            tot += kv.second;
          std::fprintf(stderr, "[pe timing] %d iterations, %.3f ms in marked phases (%.3f ms per iteration)\n", iterations, tot,
                       tot / (iterations > 0 ? iterations : 1));
          for (auto &kv : acc)
            std::fprintf(stderr, "[pe timing]   %-28s %9.3f ms  %6d x  %8.1f us each  %5.1f%%\n", kv.first.c_str(), kv.second,
                         cnt[kv.first], 1e3 * kv.second / cnt[kv.first], 100. * kv.second / tot);
        }
      acc.clear();
      cnt.clear();
      marks.clear();
      used = 0;
    }
  };
#define PE_MARK(stream, name) ::pe::PhaseTimer::get().mark(stream, name)
  // NVTX range over one C-ABI call (assemble_system, solve_time_step, ...)
  struct NvtxRange
  {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
  };
} // namespace pe
