// Job control of the C ABI (include/rtu.h): the non-blocking frame and the PNG writer that overlaps the next frame.
//
// The reference starts a frame with BeginRender() - "renderer must run in a separate thread" (viewport.cpp:36,443;
// main.cpp:66-68 detaches a thread that spawns the render threads) - the viewport then polls
// renderImage.GetNumRenderedPixels() (scene.h:585-588, viewport.cpp:390-410) and draws whatever the image holds, and
// StopRender() (main.cpp:70-72) is meant to abort it.  rtu_render_async is that contract for the device path: a host
// worker thread renders the frame in slices (groups of samples, or row blocks when there are few samples), after every
// slice the caller's image holds the mean over what is done so far and the progress counter moves; rtu_job_cancel is
// StopRender().  While a job runs its context belongs to the worker (one context per host thread, like everywhere else).
#include <atomic>
#include <cstring>
#include <thread>

#include "rtu_objects.h"

struct rtu_job {
    std::thread worker;
    std::atomic<long long> done{0};
    long long total = 0;
    std::atomic<int> cancel{0};
    std::atomic<int> finished{0};
    int status = RTU_OK;
    std::string error;
};

namespace {

struct Slice { int s0, s1, r0, r1; };

void run_frame(rtu_job *job, rtu_scene *s, rtu_params p, rtu_image out, rtu_progress_fn cb, void *user)
{
    auto fail = [&](int rc) {
        job->status = rc;
        job->error = rtu::last_error();
        job->finished.store(1);
    };
    int W, H, rc;
    if ((rc = rtu_frame_dims(s, &p, &W, &H))) return fail(rc);
    if (p.spp < 1) { rtu::set_error("rtu_render_async: bad spp"); return fail(RTU_ERR_INVALID); }
    int s_lo = 0, s_hi = p.spp, r_lo = 0, r_hi = H;
    if (p.sample_begin != 0 || p.sample_end != 0) { s_lo = p.sample_begin; s_hi = p.sample_end; }
    if (p.row_begin != 0 || p.row_end != 0) { r_lo = p.row_begin; r_hi = p.row_end; }
    if (s_lo < 0 || s_hi > p.spp || s_lo >= s_hi || r_lo < 0 || r_hi > H || r_lo >= r_hi) { rtu::set_error("rtu_render_async: bad sample / row range"); return fail(RTU_ERR_INVALID); }
    // slices: up to 16 groups of samples; frames with few samples are cut into row blocks instead (multiples of the 4-row tiles)
    std::vector<Slice> slices;
    const int ns = s_hi - s_lo;
    if (ns >= 8 || p.mode == RTU_MODE_PRIMARY) {
        const int k = p.mode == RTU_MODE_PRIMARY ? 1 : (ns < 16 ? ns : 16);
        for (int i = 0; i < k; i++) slices.push_back({s_lo + (int)((long long)ns * i / k), s_lo + (int)((long long)ns * (i + 1) / k), r_lo, r_hi});
    } else {
        const int rows = r_hi - r_lo, k = rows >= 64 ? 8 : 1;
        for (int i = 0; i < k; i++) {
            int a = r_lo + (int)((long long)rows * i / k) / 4 * 4, b = i + 1 == k ? r_hi : r_lo + (int)((long long)rows * (i + 1) / k) / 4 * 4;
            if (a < b) slices.push_back({s_lo, s_hi, a, b});
        }
    }
    const long long npix = (long long)W * H;
    job->total = npix;
    if (p.mode == RTU_MODE_PRIMARY) { // ids / z only: one call
        rc = rtu_render(s, &p, &out);
        if (rc) return fail(rc);
        job->done.store(npix);
        if (cb) cb(user, npix, npix);
        job->finished.store(1);
        return;
    }
    long long work_done = 0;
    const long long work_total = (long long)ns * (r_hi - r_lo);
    for (size_t i = 0; i < slices.size(); i++) {
        if (job->cancel.load()) { rtu::set_error("rtu_render_async: cancelled"); return fail(RTU_ERR_CANCELLED); }
        rtu_params q = p;
        q.sample_begin = slices[i].s0; q.sample_end = slices[i].s1;
        q.row_begin = slices[i].r0; q.row_end = slices[i].r1;
        const bool last = i + 1 == slices.size();
        if ((rc = rtu_render_checked(s, &q, nullptr, i == 0 ? 1 : 0, nullptr))) return fail(rc);
        work_done += (long long)(slices[i].s1 - slices[i].s0) * (slices[i].r1 - slices[i].r0);
        if (cb || last) {
            // what the viewport would draw now: the mean over the samples that are done (all of them for finished rows)
            rtu_params v = p;
            const bool by_samples = slices[i].r0 == r_lo && slices[i].r1 == r_hi;
            if (by_samples) v.spp = slices[i].s1 - s_lo;
            if (last) v.spp = p.spp; // the finished frame divides by maxSampleSize like Render() (RenderFunctions.cpp:152)
            v.sample_begin = v.sample_end = 0;
            v.row_begin = r_lo; v.row_end = slices[i].r1;
            if ((rc = rtu_resolve(s, &v, nullptr, &out))) return fail(rc);
        }
        const long long pixels = last ? npix : npix * work_done / work_total; // numRenderedPixels reaches W*H at the end (scene.h:588)
        job->done.store(pixels);
        if (cb) cb(user, pixels, npix);
    }
    job->finished.store(1);
}

} // namespace

extern "C" {

int rtu_render_async(rtu_scene *s, const rtu_params *p, const rtu_image *out, rtu_progress_fn cb, void *user, rtu_job **job_out)
{
    if (!s || !p || !out || !job_out) { rtu::set_error("rtu_render_async: null argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_render_async", [&]() -> int {
        rtu_job *job = new rtu_job;
        const rtu_params pc = *p;
        const rtu_image oc = *out;
        job->worker = std::thread([job, s, pc, oc, cb, user]() {
            try {
                run_frame(job, s, pc, oc, cb, user);
            } catch (...) {
                job->status = RTU_ERR_INVALID;
                job->error = "rtu_render_async: exception in the worker thread";
                job->finished.store(1);
            }
        });
        *job_out = job;
        return RTU_OK;
    });
}

int rtu_job_progress(const rtu_job *job, int64_t *pixels_done, int64_t *pixels_total, int32_t *finished)
{
    if (!job) { rtu::set_error("rtu_job_progress: null job"); return RTU_ERR_INVALID; }
    if (pixels_done) *pixels_done = job->done.load();
    if (pixels_total) *pixels_total = job->total;
    if (finished) *finished = job->finished.load();
    return RTU_OK;
}

void rtu_job_cancel(rtu_job *job)
{
    if (job) job->cancel.store(1);
}

int rtu_job_wait(rtu_job *job)
{
    if (!job) { rtu::set_error("rtu_job_wait: null job"); return RTU_ERR_INVALID; }
    if (job->worker.joinable()) job->worker.join();
    if (job->status) rtu::set_error(job->error);
    return job->status;
}

void rtu_job_destroy(rtu_job *job)
{
    if (!job) return;
    if (job->worker.joinable()) job->worker.join();
    delete job;
}

int rtu_write_png_async(const char *path, const uint8_t *pixels, int32_t width, int32_t height, int32_t channels, rtu_job **job_out)
{
    if (!path || !pixels || !job_out || width <= 0 || height <= 0 || (channels != 1 && channels != 3)) { rtu::set_error("rtu_write_png_async: bad argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_write_png_async", [&]() -> int {
        rtu_job *job = new rtu_job;
        // the pixels are copied: the caller's buffer is free for the next frame as soon as this returns
        auto copy = std::make_shared<std::vector<uint8_t>>(pixels, pixels + (size_t)width * height * channels);
        const std::string file = path;
        job->total = (long long)width * height;
        job->worker = std::thread([job, copy, file, width, height, channels]() {
            std::string err;
            bool ok = false;
            try {
                ok = rtu::encode_png(file.c_str(), copy->data(), width, height, channels, &err);
            } catch (...) {
                err = "rtu_write_png_async: exception in the worker thread";
            }
            if (!ok) { job->status = RTU_ERR_IO; job->error = err; }
            else job->done.store(job->total);
            job->finished.store(1);
        });
        *job_out = job;
        return RTU_OK;
    });
}

} // extern "C"
