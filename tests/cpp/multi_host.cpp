// tests/cpp/multi_host.cpp -- CPU test of the multi-device batches' host plumbing (csrc/r8b_multi.cpp): the worker
// pool that fans a call out to the shards, the error relay, thread placement, and the pinned allocator's behaviour on
// a box without a CUDA device (it must refuse, not hand out unpinned memory).  No kernel runs here.
#include "../../r8brain-free-src_b200/csrc/r8b_multi.h"

#include <sched.h>

#include <atomic>
#include <cstdio>
#include <thread>

using namespace r8bgpu;

static thread_local std::string t_err;

#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c);        \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main()
{
    // 1. every shard's job runs on that shard's own, long-lived thread; results come back by shard index
    {
        ShardPool pool(std::vector<int>(5, -1));
        CHECK(pool.size() == 5);
        std::vector<std::thread::id> first(5), again(5);
        std::vector<std::string> errs;
        auto last_error = [] { return t_err; };
        std::function<int(int)> job = [&](int s) {
            first[s] = std::this_thread::get_id();
            return 100 + s;
        };
        std::vector<int> r = pool.run_all(job, &errs, last_error);
        for (int s = 0; s < 5; s++) CHECK(r[s] == 100 + s && errs[s].empty());
        std::function<int(int)> job2 = [&](int s) {
            again[s] = std::this_thread::get_id();
            return s;
        };
        pool.run_all(job2, nullptr, last_error);
        for (int s = 0; s < 5; s++) {
            CHECK(first[s] == again[s]);
            CHECK(first[s] != std::this_thread::get_id());
            for (int u = 0; u < s; u++) CHECK(first[s] != first[u]);
        }

        // 2. a failing shard reports ITS thread-local error text; the others stay clean
        std::function<int(int)> bad = [&](int s) {
            if (s == 3) {
                t_err = "shard three refused";
                return -1;
            }
            return 7;
        };
        r = pool.run_all(bad, &errs, last_error);
        for (int s = 0; s < 5; s++) {
            CHECK(r[s] == (s == 3 ? -1 : 7));
            CHECK(errs[s] == (s == 3 ? "shard three refused" : ""));
        }

        // 3. many rounds back to back: no lost wake-up, every job of every round runs exactly once
        std::atomic<long> sum{0};
        std::function<int(int)> add = [&](int s) {
            sum += s + 1;
            return 0;
        };
        for (int it = 0; it < 20000; it++) pool.run_all(add, nullptr, last_error);
        CHECK(sum.load() == 20000L * 15);
    } // 4. the destructor joins idle workers

    // 5. placement: binding to a node keeps the thread inside the mask the process was given
    cpu_set_t before, after;
    CHECK(sched_getaffinity(0, sizeof before, &before) == 0);
    CHECK(bind_thread_to_node(-1)); // "unknown node" is a no-op
    const int nodes = numa_node_count();
    if (nodes > 0) {
        std::thread([&] {
            if (bind_thread_to_node(0)) {
                sched_getaffinity(0, sizeof after, &after);
                cpu_set_t both;
                CPU_AND(&both, &after, &before);
                if (!CPU_EQUAL(&both, &after) || CPU_COUNT(&after) == 0) printf("FAIL binding left the process mask\n");
            }
        }).join();
    }
    CHECK(!bind_thread_to_node(1 << 20)); // no such node

    // 6. without a CUDA device the pinned allocator refuses (no silent unpinned memory), and free() knows its own
    int local = 0;
    CHECK(!numa_host_free(&local));
    void* p = numa_host_alloc(1 << 20, {{0, 1 << 19, 0}, {1 << 19, 1 << 19, nodes > 1 ? 1 : 0}});
    if (p != nullptr) { // a GPU is present: the memory must be usable and releasable
        ((char*) p)[12345] = 1;
        CHECK(numa_host_free(p));
        CHECK(!numa_host_free(p));
        printf("OK (device present) nodes=%d\n", nodes);
    } else
        printf("OK (no device: allocation refused) nodes=%d\n", nodes);
    return 0;
}
