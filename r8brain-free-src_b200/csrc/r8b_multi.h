// r8b_multi.h -- host-side plumbing of multi-device batches (r8bgpu_batch_create(plan, n, -1)): one worker thread per
// device shard, NUMA placement of the threads and of the pinned host buffers they copy from / into.
//
// Channels are independent streams (one reference object per channel, example.cpp:30-67), so a batch shards over the
// GPUs of a box with no device-to-device traffic: shard s owns the contiguous channels [ch0, ch0 + n_ch) and runs the
// ordinary single-device engine on them.  What decides end-to-end throughput is the host side: every shard moves
// its own slice of the caller's buffers over its own PCIe link, and on two-socket boxes that only runs at full rate
// when the pages and the submitting thread sit on the socket the GPU hangs off.  Nothing here touches CUDA kernels.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace r8bgpu {

// NUMA node of a CUDA device (from sysfs, via its PCI bus id); -1 when unknown or the box has a single node.
int gpu_numa_node(int device);
// Restrict the calling thread to the CPUs of `node` (no-op for node < 0).
bool bind_thread_to_node(int node);
int numa_node_count();

// Page-locked host memory whose byte ranges are placed on given NUMA nodes: mmap + mbind per range + cudaHostRegister.
struct NumaRange {
    size_t offset, bytes;
    int node; // < 0: leave to first touch
};
void* numa_host_alloc(size_t bytes, const std::vector<NumaRange>& ranges);
// Releases memory from numa_host_alloc(); returns false if `p` is not one of its allocations.
bool numa_host_free(void* p);

// One long-lived worker per shard; run_all() hands every worker the same callable and waits for all of them.
class ShardPool {
public:
    explicit ShardPool(const std::vector<int>& numa_nodes);
    ~ShardPool();
    // results[s] = fn(s); fn runs on shard s's own thread
    std::vector<int> run_all(const std::function<int(int)>& fn, std::vector<std::string>* errors,
                             const std::function<std::string()>& last_error);
    int size() const { return (int) workers_.size(); }

private:
    struct Worker {
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        const std::function<int(int)>* job = nullptr;
        const std::function<std::string()>* err_fn = nullptr;
        int result = 0;
        std::string error;
        bool has_job = false, done = false, quit = false;
    };
    std::vector<Worker*> workers_;
    void loop(Worker* w, int index, int node);
};

} // namespace r8bgpu
