// Host side of the photon map: the left-balanced kd-tree of cyPhotonMap (cyPhotonMap.h:207-290) built over caller
// photons.  The tree is stored the way the reference stores it (heap order, slot 0 unused, splitting axis in the low
// two bits of plane_dirz), so a map balanced here can be handed to the reference and vice versa, and the device
// gather walks the same nodes the reference's LocatePhotons walks.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "host_scene.h"

namespace {

// What the selection moves around: the position it compares and where the photon came from (16 bytes instead of the
// 24-byte record: a third less memory traffic in the partition passes, which are all this build does).
struct Item {
    float position[3];
    uint32_t src; // index into the caller's array
};

struct Balancer {
    Item *work;           // 1-based; partitioned in place
    const rtu_photon *in; // the caller's photons (0-based)
    rtu_photon *out;      // 1-based heap order
    void place(int slot, const Item &it, int axis_bits)
    {
        out[slot] = in[it.src];
        if (axis_bits >= 0) out[slot].plane_dirz = (uint8_t)((out[slot].plane_dirz & 0x8) | axis_bits);
    }

    // Position of the median that leaves a complete left subtree (cyPhotonMap.h:233-241)
    static int median_of(int start, int end)
    {
        const int count = end - start + 1;
        int m = 1;
        while (4 * m <= count) m += m;
        if (3 * m <= count) return 2 * m + start - 1;
        return end - m + 1;
    }

    // Hoare-style selection around the value of the right-most element until `median` is in place (:252-267)
    void select(int axis, int start, int end, int median)
    {
        if (team && end - start + 1 >= PAR_MIN) select_parallel(axis, start, end, median);
        else select_from(axis, start, end, median);
    }

    // ---- the top levels.  A range of hundreds of thousands of photons is one selection on one thread in the reference, and the
    // first three levels of a 10^6-photon map were two thirds of the build.  One PASS of that selection (Hoare partition
    // around the last element) has a closed form: with I = the positions, ascending, where the up-scan stops (a >= pivot;
    // the pivot's own slot ends it) and J = the positions, descending, where the down-scan stops (a <= pivot, or the range's
    // first slot, which stops it unconditionally), the pass swaps I[k] with J[k] for every k < K, K = the first k with
    // I[k] >= J[k] - until then the scans only see slots no swap has touched - and the pivot then goes to min(I[K], J[K-1]):
    // the up-scan's last stop is its next original one or the slot the last swap has just filled from the other side.  Both lists are
    // stream compactions, K is a bisection (I ascends, J descends), the swaps are independent: every step runs on all
    // threads, and the array after the pass is the reference's, element for element.
    struct Team {
        unsigned n = 1;
        std::vector<std::thread> th;
        std::atomic<unsigned> arrived{0}, phase{0};
        std::atomic<int> job{0}; // 0 idle, 1 run, -1 quit
        std::function<void(unsigned)> fn;
        void barrier()
        {
            const unsigned p = phase.load(std::memory_order_acquire);
            if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
                arrived.store(0, std::memory_order_relaxed);
                phase.store(p + 1, std::memory_order_release);
            } else {
                while (phase.load(std::memory_order_acquire) == p) std::this_thread::yield();
            }
        }
        // runs f(t) on every member (the caller is member 0) and returns when all are done
        void run(const std::function<void(unsigned)> &f)
        {
            fn = f;
            barrier(); // release the workers
            fn(0);
            barrier(); // everybody done
        }
        void start(unsigned members)
        {
            n = members;
            for (unsigned t = 1; t < n; t++)
                th.emplace_back([this, t]() {
                    for (;;) {
                        barrier();
                        if (job.load(std::memory_order_acquire) < 0) return;
                        fn(t);
                        barrier();
                    }
                });
        }
        void stop()
        {
            if (th.empty()) return;
            job.store(-1, std::memory_order_release);
            barrier();
            for (auto &t : th) t.join();
            th.clear();
        }
    };
    Team *team = nullptr;
    struct Deferred { int index, start, end; float lo[3], hi[3]; };
    std::vector<Deferred> later;      // subtrees below PAR_MIN met while the team was busy with the top levels
    uint32_t *ipos = nullptr, *jpos = nullptr; // the two stop lists of a pass
    static constexpr int PAR_MIN = 1 << 17; // ranges at least this long are selected by the whole team

    void select_parallel(int axis, int start, int end, int median)
    {
        const unsigned T = team->n;
        std::vector<uint32_t> cnt_i(T + 1), cnt_j(T + 1);
        int left = start, right = end;
        while (right > left) {
            if (right - left + 1 < PAR_MIN / 4) { select_from(axis, left, right, median); return; } // the tail of the selection: one thread
            const float pivot = work[right].position[axis];
            const int n = right - left + 1; // slots left .. right; the up-scan may stop on any of them, the down-scan on left .. right-1
            auto chunk = [&](unsigned t, int &a, int &b) { a = left + (int)((long long)n * t / T); b = left + (int)((long long)n * (t + 1) / T); };
            team->run([&](unsigned t) {
                int a, b;
                chunk(t, a, b);
                uint32_t ci = 0, cj = 0;
                for (int p = a; p < b; p++) {
                    const float v = work[p].position[axis];
                    ci += !(v < pivot);                         // where `while (a[++i] < pivot)` stops (the pivot's own slot at the latest)
                    cj += p < right && (!(v > pivot) || p == left); // where `while (a[--j] > pivot && j > left)` stops
                }
                cnt_i[t + 1] = ci;
                cnt_j[t + 1] = cj;
            });
            cnt_i[0] = cnt_j[0] = 0;
            for (unsigned t = 0; t < T; t++) { cnt_i[t + 1] += cnt_i[t]; cnt_j[t + 1] += cnt_j[t]; }
            const uint32_t ni = cnt_i[T], nj = cnt_j[T];
            team->run([&](unsigned t) {
                int a, b;
                chunk(t, a, b);
                uint32_t oi = cnt_i[t], oj = cnt_j[t];
                for (int p = a; p < b; p++) {
                    const float v = work[p].position[axis];
                    if (!(v < pivot)) ipos[oi++] = (uint32_t)p;
                    if (p < right && (!(v > pivot) || p == left)) jpos[nj - 1 - oj++] = (uint32_t)p; // descending
                }
            });
            // K = first k with I[k] >= J[k]
            uint32_t lo = 0, hi = ni < nj ? ni : nj;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) / 2;
                if (ipos[mid] < jpos[mid]) lo = mid + 1; else hi = mid;
            }
            const uint32_t K = lo;
            team->run([&](unsigned t) {
                const uint32_t a = (uint32_t)((unsigned long long)K * t / T), b = (uint32_t)((unsigned long long)K * (t + 1) / T);
                for (uint32_t k = a; k < b; k++) std::swap(work[ipos[k]], work[jpos[k]]);
            });
            // where the up-scan stops for the last time: its next stop of the original array (I[K] exists: the pivot's slot is
            // the last entry of I and lies above every entry of J) - or the slot the last swap filled from the other side
            int i = (int)ipos[K];
            if (K > 0 && (int)jpos[K - 1] < i) i = (int)jpos[K - 1];
            std::swap(work[i], work[right]);
            if (i >= median) right = i - 1;
            if (i <= median) left = i + 1;
        }
    }

    // the reference's loop from an intermediate state (left, right) of a selection
    void select_from(int axis, int left, int right, int median)
    {
        while (right > left) {
            const float pivot = work[right].position[axis];
            int i = left - 1, j = right;
            for (;;) {
                while (work[++i].position[axis] < pivot) {}
                while (work[--j].position[axis] > pivot && j > left) {}
                if (i >= j) break;
                std::swap(work[i], work[j]);
            }
            std::swap(work[i], work[right]);
            if (i >= median) right = i - 1;
            if (i <= median) left = i + 1;
        }
    }

    // The two sides of a median are disjoint ranges of `work` and disjoint heap slots of `out`, so below the top few
    // levels the recursion runs on separate threads; the result is the sequential one.
    void segment(const float lo[3], const float hi[3], int index, int start, int end, int fork_levels)
    {
        const int median = median_of(start, end);
        const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        int axis = 2; // widest extent; ties fall through exactly like :244-248
        if (ex > ey) {
            if (ex > ez) axis = 0;
        } else if (ey > ez) {
            axis = 1;
        }
        select(axis, start, end, median);
        place(index, work[median], axis);
        const float split = work[median].position[axis];
        float h2[3] = {hi[0], hi[1], hi[2]}, l2[3] = {lo[0], lo[1], lo[2]};
        h2[axis] = split;
        l2[axis] = split;
        const bool left_rec = median > start && start < median - 1, right_rec = median < end && median + 1 < end;
        if (median > start && !left_rec) place(2 * index, work[start], -1);
        if (median < end && !right_rec) place(2 * index + 1, work[end], -1);
        if (team && end - start + 1 >= PAR_MIN) {
            // both sides one after the other while their selections still use the whole team; sides below PAR_MIN are put off
            // until the team is idle (`later`), then run side by side
            for (int side = 0; side < 2; side++) {
                const bool rec = side == 0 ? left_rec : right_rec;
                if (!rec) continue;
                Deferred d;
                d.index = side == 0 ? 2 * index : 2 * index + 1;
                d.start = side == 0 ? start : median + 1;
                d.end = side == 0 ? median - 1 : end;
                for (int k = 0; k < 3; k++) { d.lo[k] = side == 0 ? lo[k] : l2[k]; d.hi[k] = side == 0 ? h2[k] : hi[k]; }
                if (d.end - d.start + 1 >= PAR_MIN) segment(d.lo, d.hi, d.index, d.start, d.end, fork_levels);
                else later.push_back(d);
            }
            return;
        }
        if (left_rec && right_rec && fork_levels > 0 && end - start > 4096) {
            std::thread t([&]() { segment(lo, h2, 2 * index, start, median - 1, fork_levels - 1); });
            segment(l2, hi, 2 * index + 1, median + 1, end, fork_levels - 1);
            t.join();
            return;
        }
        if (left_rec) segment(lo, h2, 2 * index, start, median - 1, fork_levels);
        if (right_rec) segment(l2, hi, 2 * index + 1, median + 1, end, fork_levels);
    }
};

} // namespace

extern "C" int rtu_host_balance_photons(const rtu_photon *in, uint32_t n, rtu_photon *out)
{
    if (!out || (n && !in)) { rtu::set_error("rtu_host_balance_photons: null argument"); return RTU_ERR_INVALID; }
    if (n >= (1u << 30)) { rtu::set_error("rtu_host_balance_photons: too many photons"); return RTU_ERR_INVALID; }
    std::memset(out, 0, sizeof(rtu_photon)); // slot 0; every slot 1..n is assigned exactly once by the recursion
    if (n == 0) return RTU_OK;
    return rtu::guarded("rtu_host_balance_photons", [&]() -> int {
    Balancer b;
    std::unique_ptr<Item[]> work(new Item[(size_t)n + 1]); // not value-initialised
    b.work = work.get();
    std::memset(&b.work[0], 0, sizeof(Item));
    b.in = in;
    b.out = out;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    // copy + bounding box in slices (min / max do not depend on the order); the reference seeds the box with the zeroed
    // slot 0 of its vector (:212-213): the origin is always inside
    const unsigned slices = n > (1u << 16) ? (hw < 16 ? hw : 16) : 1;
    std::vector<float> slo(3 * slices, 0.f), shi(3 * slices, 0.f);
    {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < slices; t++) {
            auto job = [&, t]() {
                const size_t a = 1 + (size_t)n * t / slices, e = 1 + (size_t)n * (t + 1) / slices;
                float l[3] = {0, 0, 0}, h[3] = {0, 0, 0};
                for (size_t i = a; i < e; i++)
                    for (int k = 0; k < 3; k++) {
                        const float v = in[i - 1].position[k];
                        b.work[i].position[k] = v;
                        b.work[i].src = (uint32_t)(i - 1);
                        if (l[k] > v) l[k] = v;
                        if (h[k] < v) h[k] = v;
                    }
                for (int k = 0; k < 3; k++) { slo[3 * t + k] = l[k]; shi[3 * t + k] = h[k]; }
            };
            if (t + 1 < slices) pool.emplace_back(job); else job();
        }
        for (auto &th : pool) th.join();
    }
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (unsigned t = 0; t < slices; t++)
        for (int k = 0; k < 3; k++) {
            if (lo[k] > slo[3 * t + k]) lo[k] = slo[3 * t + k];
            if (hi[k] < shi[3 * t + k]) hi[k] = shi[3 * t + k];
        }
    int fork_levels = 0;
    while ((1u << (fork_levels + 1)) <= hw && fork_levels < 6) fork_levels++;
    Balancer::Team team;
    std::unique_ptr<uint32_t[]> lists;
    if ((int)n >= Balancer::PAR_MIN && hw > 1 && !getenv("RTU_PHOTON_SERIAL_TOP")) {
        team.start(hw < 32 ? hw : 32);
        b.team = &team;
        lists.reset(new uint32_t[2 * ((size_t)n + 1)]); // not value-initialised
        b.ipos = lists.get();
        b.jpos = lists.get() + (size_t)n + 1;
    }
    b.segment(lo, hi, 1, 1, (int)n, fork_levels);
    team.stop();
    b.team = nullptr;
    if (!b.later.empty()) {
        // the put-off subtrees side by side (each forks further like the plain recursion does)
        int fl = 0;
        while ((b.later.size() << (fl + 1)) <= (size_t)hw && fl < 6) fl++;
        std::vector<std::thread> pool;
        for (size_t k = 1; k < b.later.size(); k++)
            pool.emplace_back([&, k]() { const auto &d = b.later[k]; b.segment(d.lo, d.hi, d.index, d.start, d.end, fl); });
        { const auto &d = b.later[0]; b.segment(d.lo, d.hi, d.index, d.start, d.end, fl); }
        for (auto &th : pool) th.join();
    }
    return RTU_OK;
    });
}

extern "C" void rtu_photon_params_default(rtu_photon_params *p)
{
    if (!p) return;
    p->map_size = 1000000; // RenderFunctions.cpp:32-36
    p->max_bounce = 10;
    p->est_radius = 1.0f;
    p->ellipticity = 0.5f;
    p->seed = 0;
}
