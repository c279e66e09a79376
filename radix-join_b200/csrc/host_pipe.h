// Host side of the page-pointer path: a persistent worker pool and rings of pinned staging buffers.
//
// The contest's ColumnarTable is a list of individually `new`-ed 8 KB pages per column
// (reference include/plan.h:54-68,95-99), so neither the inputs nor the result can be handed to the
// DMA engines directly: input pages are GATHERED into pinned buffers by worker threads and copied to
// the device buffer by buffer, result pages arrive in pinned buffers and are SCATTERED into freshly
// allocated pages.  Everything here exists so that those memcpys, the two DMA directions and the
// kernels all run at the same time (rj_execute_pages, engine.cu):
//
//   ThreadPool   fixed set of workers created with the context (spawning threads per chunk cost more
//                than the copies they ran); every worker binds the context's device once
//   TaskGroup    completion counter of a batch of tasks; the LAST finisher runs `on_complete`
//                (used to record the "window uploaded" CUDA event from whichever worker ends last)
//   PinnedRing   blocking pool of fixed-size pinned buffers, allocated lazily up to a cap and kept for
//                the life of the context; each buffer carries the CUDA event of the last DMA that used it
//   EventWaiter  one thread that waits for D2H events IN ORDER and only then hands the scatter task to
//                the pool, so no worker ever sleeps on an event while gather tasks are queued
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <stdexcept>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace rj {

class ThreadPool {
public:
    ThreadPool(int n_threads, int device): device_(device) {
        if (n_threads < 1) n_threads = 1;
        for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { run(); });
    }
    ~ThreadPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t: workers_) t.join();
    }
    int  size() const { return static_cast<int>(workers_.size()); }
    void submit(std::function<void()> fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            queue_.push_back(std::move(fn));
        }
        cv_.notify_one();
    }

private:
    void run() {
        cudaSetDevice(device_);
        for (;;) {
            std::function<void()> fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !queue_.empty(); });
                if (queue_.empty()) return; // stop_ and drained
                fn = std::move(queue_.front());
                queue_.pop_front();
            }
            fn();
        }
    }
    int                               device_;
    std::mutex                        mu_;
    std::condition_variable           cv_;
    std::deque<std::function<void()>> queue_;
    std::vector<std::thread>          workers_;
    bool                              stop_ = false;
};

// A batch of tasks.  add() before submitting, done() at the end of each task (fail() instead when it
// threw); wait() blocks until the count is back to zero and rethrows the first failure.
class TaskGroup {
public:
    std::function<void()> on_complete; // run once by the task that brings the count to zero
    void add(int64_t n = 1) { pending_.fetch_add(n, std::memory_order_relaxed); }
    void done() {
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) finish();
    }
    void fail(std::exception_ptr e) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!err_) err_ = e;
        }
        done();
    }
    // guard of the submitting thread: hold one count while the tasks are being submitted so the group
    // cannot complete early; call done() (via seal) when every task is in the queue
    void open() { add(1); }
    void seal() { done(); }
    bool complete() {
        std::lock_guard<std::mutex> lk(mu_);
        return finished_;
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return finished_; });
        if (err_) std::rethrow_exception(err_);
    }
    bool failed() {
        std::lock_guard<std::mutex> lk(mu_);
        return static_cast<bool>(err_);
    }

private:
    void finish() {
        if (on_complete) {
            try {
                on_complete();
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu_);
                if (!err_) err_ = std::current_exception();
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            finished_ = true;
        }
        cv_.notify_all();
    }
    std::atomic<int64_t>    pending_{0};
    std::mutex              mu_;
    std::condition_variable cv_;
    std::exception_ptr      err_;
    bool                    finished_ = false;
};

struct PinnedBuf {
    uint8_t*    p = nullptr;
    cudaEvent_t ev = nullptr; // last DMA that read or wrote the buffer
};

class PinnedRing {
public:
    PinnedRing(size_t buf_bytes, int cap): bytes_(buf_bytes), cap_(cap) {}
    ~PinnedRing() {
        for (auto* b: all_) {
            if (b->p) cudaFreeHost(b->p);
            if (b->ev) cudaEventDestroy(b->ev);
            delete b;
        }
    }
    size_t buf_bytes() const { return bytes_; }
    // a buffer nobody uses (its last DMA has completed); blocks while `cap` buffers are out
    PinnedBuf* acquire() {
        PinnedBuf* b = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu_);
            for (;;) {
                if (!free_.empty()) {
                    b = free_.front();
                    free_.pop_front();
                    break;
                }
                if (static_cast<int>(all_.size()) < cap_) {
                    b = new PinnedBuf;
                    all_.push_back(b);
                    break;
                }
                cv_.wait(lk);
            }
        }
        if (!b->p) {
            if (cudaMallocHost(reinterpret_cast<void**>(&b->p), bytes_) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->ev, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                if (b->p) cudaFreeHost(b->p);
                b->p = nullptr;
                b->ev = nullptr;
                release(b); // the next taker tries again
                throw std::runtime_error("pinned staging buffer: cudaMallocHost failed");
            }
        } else {
            cudaEventSynchronize(b->ev);
        }
        return b;
    }
    void release(PinnedBuf* b) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            free_.push_back(b);
        }
        cv_.notify_one();
    }

private:
    size_t                  bytes_;
    int                     cap_;
    std::mutex              mu_;
    std::condition_variable cv_;
    std::deque<PinnedBuf*>  free_;
    std::vector<PinnedBuf*> all_;
};

// Waits for CUDA events in submission order, then forwards the attached task to the pool.
class EventWaiter {
public:
    EventWaiter(ThreadPool* pool, int device): pool_(pool), device_(device), thread_([this] { run(); }) {}
    ~EventWaiter() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        thread_.join();
    }
    void after(cudaEvent_t ev, std::function<void()> fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            queue_.emplace_back(ev, std::move(fn));
        }
        cv_.notify_one();
    }

private:
    void run() {
        cudaSetDevice(device_);
        for (;;) {
            std::pair<cudaEvent_t, std::function<void()>> item;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !queue_.empty(); });
                if (queue_.empty()) return;
                item = std::move(queue_.front());
                queue_.pop_front();
            }
            cudaEventSynchronize(item.first); // a failed copy surfaces on the stream; the task still runs and releases its buffer
            pool_->submit(std::move(item.second));
        }
    }
    ThreadPool*                                                  pool_;
    int                                                          device_;
    std::mutex                                                   mu_;
    std::condition_variable                                      cv_;
    std::deque<std::pair<cudaEvent_t, std::function<void()>>>    queue_;
    bool                                                         stop_ = false;
    std::thread                                                  thread_;
};

struct HostPipe {
    static constexpr size_t kBufBytes = size_t(4) << 20; // 512 pages per staging buffer
    static constexpr int    kUpCap    = 64;              // 256 MiB of upload staging
    static constexpr int    kDownCap  = 96;              // 384 MiB of download staging
    ThreadPool  pool;
    PinnedRing  up, down;
    EventWaiter waiter;
    HostPipe(int n_threads, int device, int cap = 0)
        : pool(n_threads, device), up(kBufBytes, cap > 0 ? cap : kUpCap), down(kBufBytes, cap > 0 ? cap : kDownCap), waiter(&pool, device) {}
};

} // namespace rj
