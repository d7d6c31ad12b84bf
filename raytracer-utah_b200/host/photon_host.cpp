// Host side of the photon map: the left-balanced kd-tree of cyPhotonMap (cyPhotonMap.h:207-290) built over caller
// photons.  The tree is stored the way the reference stores it (heap order, slot 0 unused, splitting axis in the low
// two bits of plane_dirz), so a map balanced here can be handed to the reference and vice versa, and the device
// gather walks the same nodes the reference's LocatePhotons walks.
#include <cstring>
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
        int left = start, right = end;
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
    b.segment(lo, hi, 1, 1, (int)n, fork_levels);
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
