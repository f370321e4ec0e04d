/*
 * sdft_host_copy.hpp -- host-side helper of libsdft_b200.so: a small pool of threads that moves tiles between
 * the library's pinned staging buffers and PAGEABLE caller memory (malloc, NumPy) while the next tile is in
 * flight over PCIe.  Pure host C++, no CUDA.
 */
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace sdftb200
{

inline size_t host_env_size(const char* name, size_t fallback)
{
  const char* v = getenv(name);
  if (!v || !*v) return fallback;
  char* end = nullptr;
  const unsigned long long x = strtoull(v, &end, 10);
  return (end && end != v) ? (size_t)x : fallback;
}

/* ------------------------------------------------------------------------------------------------
 * PAGEABLE caller buffers.  The reference's callers hand over malloc'ed / NumPy memory.  A cudaMemcpy
 * from or to pageable memory is staged by the driver on ONE thread (and blocks the calling thread), so
 * rows would leave the GPU at a fraction of the PCIe rate.  Instead the library DMAs tiles into its own
 * pinned staging buffers and moves them to the caller's pages with a few host threads, the copy of tile
 * i overlapping the DMA of tile i+1.
 * ---------------------------------------------------------------------------------------------- */
struct CopySeg { void* dst; const void* src; size_t bytes; };

class HostCopier
{
public:
  static HostCopier& get()
  {
    static HostCopier* instance = new HostCopier();   // never destroyed: its threads outlive main()
    return *instance;
  }
  /* copies all segments, cut into slices, on the pool plus the calling thread; returns when done */
  void run(const std::vector<CopySeg>& segs)
  {
    const size_t slice = (size_t)4 << 20;
    std::vector<CopySeg> work;
    for (const CopySeg& s : segs)
      for (size_t off = 0; off < s.bytes; off += slice)
        work.push_back({ (char*)s.dst + off, (const char*)s.src + off, (s.bytes - off < slice) ? s.bytes - off : slice });
    if (work.empty()) return;
    size_t total = 0;
    for (const CopySeg& w : work) total += w.bytes;
    if (total < ((size_t)1 << 20))
    {
      /* waking the pool costs more than copying a few hundred KiB */
      for (const CopySeg& w : work) memcpy(w.dst, w.src, w.bytes);
      return;
    }
    std::lock_guard<std::mutex> one_caller(run_mutex_);   // plans on different threads take turns
    {
      std::unique_lock<std::mutex> lock(mutex_);
      work_ = &work;
      next_.store(0);
      pending_ = work.size();
      ++generation_;
    }
    wake_.notify_all();
    drain(&work);
    /* `work` lives on this stack frame: wait until every slice is copied AND every pool thread that
     * picked this job up has let go of it */
    std::unique_lock<std::mutex> lock(mutex_);
    work_ = nullptr;                      // threads waking up late find nothing to do
    done_.wait(lock, [&] { return pending_ == 0 && active_ == 0; });
  }

private:
  HostCopier()
  {
    unsigned n = std::thread::hardware_concurrency();
    n = (n > 2) ? n / 2 : 1;
    if (n > 8) n = 8;
    n = (unsigned)host_env_size("SDFT_B200_COPY_THREADS", n);
    for (unsigned i = 1; i < n; ++i) std::thread([this] { loop(); }).detach();   // the caller is thread 0
  }
  void drain(const std::vector<CopySeg>* work)
  {
    size_t finished = 0;
    while (true)
    {
      const size_t i = next_.fetch_add(1);
      if (i >= work->size()) break;
      memcpy((*work)[i].dst, (*work)[i].src, (*work)[i].bytes);
      ++finished;
    }
    if (finished)
    {
      std::unique_lock<std::mutex> lock(mutex_);
      pending_ -= finished;
      if (pending_ == 0) done_.notify_all();
    }
  }
  void loop()
  {
    unsigned long long seen = 0;
    while (true)
    {
      const std::vector<CopySeg>* work = nullptr;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
        work = work_;
        if (work) ++active_;
      }
      if (work)
      {
        drain(work);
        std::unique_lock<std::mutex> lock(mutex_);
        if (--active_ == 0 && pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex mutex_, run_mutex_;
  std::condition_variable wake_, done_;
  const std::vector<CopySeg>* work_ = nullptr;
  std::atomic<size_t> next_{ 0 };
  size_t pending_ = 0;
  size_t active_ = 0;                    // pool threads currently holding the job
  unsigned long long generation_ = 0;
};

}  // namespace sdftb200
