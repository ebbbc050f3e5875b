// r8b_multi.cpp -- see r8b_multi.h.
#include "r8b_multi.h"

#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>

namespace r8bgpu {

namespace {

bool read_file(const std::string& path, std::string& out)
{
    FILE* f = fopen(path.c_str(), "r");
    if (f == nullptr) return false;
    char buf[4096];
    const size_t n = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    buf[n] = 0;
    out = buf;
    return true;
}

// "0-31,64-95" -> cpu_set_t
bool parse_cpulist(const std::string& s, cpu_set_t& set)
{
    CPU_ZERO(&set);
    const char* p = s.c_str();
    bool any = false;
    while (*p) {
        while (*p && !isdigit((unsigned char) *p)) p++;
        if (!*p) break;
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (*e == '-') b = strtol(e + 1, &e, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET((int) c, &set);
            any = true;
        }
        p = e;
    }
    return any;
}

std::mutex g_alloc_m;
std::map<void*, size_t> g_allocs;

} // namespace

int numa_node_count()
{
    int n = 0;
    std::string s;
    while (read_file("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist", s)) n++;
    return n;
}

int gpu_numa_node(int device)
{
    if (numa_node_count() < 2) return -1;
    char bus[64] = {};
    if (cudaDeviceGetPCIBusId(bus, (int) sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* c = bus; *c; c++) *c = (char) tolower((unsigned char) *c);
    std::string s;
    if (!read_file(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", s)) return -1;
    const int node = atoi(s.c_str());
    return node >= 0 ? node : -1;
}

bool bind_thread_to_node(int node)
{
    if (node < 0) return true;
    std::string s;
    if (!read_file("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", s)) return false;
    cpu_set_t want, have;
    if (!parse_cpulist(s, want)) return false;
    // stay inside the mask the process was given (containers, taskset)
    if (sched_getaffinity(0, sizeof have, &have) == 0) {
        cpu_set_t both;
        CPU_AND(&both, &want, &have);
        if (CPU_COUNT(&both) == 0) return false;
        want = both;
    }
    return sched_setaffinity(0, sizeof want, &want) == 0;
}

void* numa_host_alloc(size_t bytes, const std::vector<NumaRange>& ranges)
{
    const size_t page = (size_t) sysconf(_SC_PAGESIZE);
    const size_t total = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
#ifdef MADV_HUGEPAGE
    // transparent huge pages, where the system grants them: 512x fewer IOMMU / page-table entries for the DMA engines
    if (getenv("R8BGPU_NO_HUGEPAGES") == nullptr) madvise(p, total, MADV_HUGEPAGE);
#endif
#ifdef SYS_mbind
    for (const NumaRange& r : ranges) {
        if (r.node < 0 || r.node >= 64 || r.bytes == 0) continue;
        // whole pages inside the range (a page shared by two ranges stays with first touch)
        const size_t a = (r.offset + page - 1) / page * page, b = (r.offset + r.bytes) / page * page;
        if (b <= a) continue;
        const unsigned long mask = 1ul << r.node;
        // MPOL_BIND = 2; failure (no permission, node offline) leaves the default policy: still correct, only slower
        syscall(SYS_mbind, (char*) p + a, b - a, 2, &mask, sizeof(mask) * 8 + 1, 0);
    }
#else
    (void) ranges;
#endif
    memset(p, 0, total); // fault the pages in under the policy set above, before they are pinned
    if (cudaHostRegister(p, total, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, total);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_alloc_m);
    g_allocs[p] = total;
    return p;
}

bool numa_host_free(void* p)
{
    size_t total = 0;
    {
        std::lock_guard<std::mutex> lk(g_alloc_m);
        auto it = g_allocs.find(p);
        if (it == g_allocs.end()) return false;
        total = it->second;
        g_allocs.erase(it);
    }
    cudaHostUnregister(p);
    munmap(p, total);
    return true;
}

ShardPool::ShardPool(const std::vector<int>& numa_nodes)
{
    for (size_t i = 0; i < numa_nodes.size(); i++) {
        Worker* w = new Worker;
        workers_.push_back(w);
        w->th = std::thread(&ShardPool::loop, this, w, (int) i, numa_nodes[i]);
    }
}

ShardPool::~ShardPool()
{
    for (Worker* w : workers_) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        w->th.join();
        delete w;
    }
}

void ShardPool::loop(Worker* w, int index, int node)
{
    bind_thread_to_node(node);
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        const std::function<int(int)>* job = w->job;
        lk.unlock();
        const int r = (*job)(index);
        std::string e;
        if (r < 0 && w->err_fn) e = (*w->err_fn)(); // the engine's error text is thread-local: fetch it on this thread
        lk.lock();
        w->result = r;
        w->error = e;
        w->has_job = false;
        w->done = true;
        w->cv.notify_all();
    }
}

std::vector<int> ShardPool::run_all(const std::function<int(int)>& fn, std::vector<std::string>* errors,
                                    const std::function<std::string()>& last_error)
{
    for (Worker* w : workers_) {
        std::lock_guard<std::mutex> lk(w->m);
        w->job = &fn;
        w->err_fn = &last_error;
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    std::vector<int> res(workers_.size());
    if (errors) errors->assign(workers_.size(), std::string());
    for (size_t i = 0; i < workers_.size(); i++) {
        Worker* w = workers_[i];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        res[i] = w->result;
        if (errors) (*errors)[i] = w->error;
    }
    return res;
}

} // namespace r8bgpu
