// Multi-GPU side of the C ABI (include/rtu.h): one process (or host thread) per GPU, each with its own
// rtu_context; the path shards by independent units, so the only data-path communication is the step that brings the
// partial images together on one rank:
//
//   rtu_reduce_resolve   spp slices (Partitioning B, the GI mode): every rank holds sums over ITS samples of ALL pixels.
//                        pack RGB planes (the float4 accumulator's .w lane carries nothing: 25 MB instead of 33 MB at
//                        1080p) -> ONE ncclReduce(sum, FP32) onto the root, on the context's stream, directly behind the
//                        frame's last kernel -> the root's resolve kernel reads the reduced planes -> device-to-host copies.
//                        No host synchronisation between render, reduce and resolve.
//   rtu_gather_resolve   row ranges (Partitioning A, deterministic Whitted frames): every rank resolves its own rows and
//                        sends the finished RGB8 / float rows to the root (grouped ncclSend / ncclRecv): 6 MB instead of 25.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries NCCL - e.g. torch's bundled copy - shares
// it, a single-GPU user never loads it, and the library has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "rtu_objects.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char *names[] = {getenv("RTU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char *n) { void *p = dlsym(api.lib, n); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    if (!api.error.empty()) { rtu::set_error(api.error); return nullptr; }
    return &api;
}

#define NC(call)                                                                                           \
    do {                                                                                                   \
        ncclResult_t r_ = (call);                                                                          \
        if (r_ != ncclSuccess) {                                                                           \
            rtu::set_error(std::string(#call) + ": " + (N->GetErrorString ? N->GetErrorString(r_) : "NCCL error")); \
            return RTU_ERR_CUDA;                                                                           \
        }                                                                                                  \
    } while (0)

} // namespace

struct rtu_comm {
    rtu_context *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    float *planes = nullptr;      // 3 x npix RGB sums: send buffer of the reduce, and its receive buffer on the root
    size_t planes_n = 0;
    int *d_rows = nullptr;        // world x 2: every rank's row range (gather mode)
    int *h_rows = nullptr;        // page-locked copy
};

extern "C" {

int rtu_comm_unique_id(uint8_t id[RTU_COMM_ID_BYTES])
{
    if (!id) { rtu::set_error("rtu_comm_unique_id: null argument"); return RTU_ERR_INVALID; }
    static_assert(sizeof(ncclUniqueId) == RTU_COMM_ID_BYTES, "ncclUniqueId size");
    NcclApi *N = nccl_api();
    if (!N) return RTU_ERR_UNSUPPORTED;
    ncclUniqueId u;
    NC(N->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return RTU_OK;
}

int rtu_comm_create(rtu_context *ctx, const uint8_t id[RTU_COMM_ID_BYTES], int32_t rank, int32_t world, rtu_comm **out)
{
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) { rtu::set_error("rtu_comm_create: bad argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_comm_create", [&]() -> int {
        NcclApi *N = nccl_api();
        if (!N) return RTU_ERR_UNSUPPORTED;
        CU(cudaSetDevice(ctx->device));
        std::unique_ptr<rtu_comm> c(new rtu_comm);
        c->ctx = ctx;
        c->rank = rank;
        c->world = world;
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        NC(N->CommInitRank(&c->comm, world, u, rank));
        cudaError_t e = cudaMalloc((void **)&c->d_rows, sizeof(int) * 2 * (size_t)world);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->h_rows, sizeof(int) * 2 * (size_t)world, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            rtu_comm *raw = c.release();
            rtu_comm_destroy(raw);
            CU(e);
        }
        *out = c.release();
        return RTU_OK;
    });
}

void rtu_comm_destroy(rtu_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    NcclApi *N = nccl_api();
    if (N && c->comm) N->CommDestroy(c->comm);
    if (c->planes) cudaFree(c->planes);
    if (c->d_rows) cudaFree(c->d_rows);
    if (c->h_rows) cudaFreeHost(c->h_rows);
    delete c;
}

static int ensure_planes(rtu_comm *m, size_t npix)
{
    if (3 * npix <= m->planes_n) return RTU_OK;
    CU(cudaStreamSynchronize(m->ctx->stream));
    if (m->planes) cudaFree(m->planes);
    m->planes = nullptr;
    m->planes_n = 0;
    CU(cudaMalloc((void **)&m->planes, 3 * npix * sizeof(float)));
    m->planes_n = 3 * npix;
    return RTU_OK;
}

int rtu_reduce_resolve(rtu_scene *s, rtu_comm *m, const rtu_params *p, const float *d_accum, int32_t root, rtu_image *out)
{
    if (!s || !m || !p || root < 0 || root >= m->world) { rtu::set_error("rtu_reduce_resolve: bad argument"); return RTU_ERR_INVALID; }
    if (m->ctx != s->ctx) { rtu::set_error("rtu_reduce_resolve: the communicator belongs to another context"); return RTU_ERR_INVALID; }
    if (m->rank == root && !out) { rtu::set_error("rtu_reduce_resolve: the root needs output buffers"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_reduce_resolve", [&]() -> int {
        NcclApi *N = nccl_api();
        if (!N) return RTU_ERR_UNSUPPORTED;
        rtu_context *c = s->ctx;
        CU(cudaSetDevice(c->device));
        int W, H, rc;
        if ((rc = rtu_frame_dims(s, p, &W, &H))) return rc;
        if (p->spp < 1) { rtu::set_error("rtu_reduce_resolve: spp must be at least 1"); return RTU_ERR_INVALID; }
        const size_t npix = (size_t)W * H;
        const float4 *accum = d_accum ? (const float4 *)d_accum : c->fb.accum;
        if (!accum || (!d_accum && npix > c->fb.accum_n)) { rtu::set_error("rtu_reduce_resolve: nothing rendered yet"); return RTU_ERR_INVALID; }
        if ((rc = ensure_planes(m, npix))) return rc;
        launch_pack_rgb(c->stream, accum, npix, m->planes);
        NC(N->Reduce(m->planes, m->planes, 3 * npix, ncclFloat, ncclSum, root, m->comm, c->stream));
        if (m->rank != root) return RTU_OK; // enqueued; this rank's host goes on (its next frame queues up behind the reduce)
        if (!out->rgb && !out->rgb8 && !out->z && !out->z8 && !out->node_id && !out->face_id) return RTU_OK; // reduce only
        if ((rc = rtu_ensure_image(s, npix))) return rc;
        if (out->rgb || out->rgb8) {
            launch_resolve_planes(c->stream, m->planes, npix, p->spp, out->rgb ? c->fb.d_rgb : nullptr, out->rgb8 ? c->fb.d_rgb8 : nullptr);
            if (out->rgb) CU(cudaMemcpyAsync(out->rgb, c->fb.d_rgb, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            if (out->rgb8) CU(cudaMemcpyAsync(out->rgb8, c->fb.d_rgb8, npix * 3, cudaMemcpyDeviceToHost, c->stream));
        }
        rtu_image rest = *out;
        rest.rgb = nullptr;
        rest.rgb8 = nullptr;
        if (rest.z || rest.z8 || rest.node_id || rest.face_id)
            if ((rc = rtu_resolve_enqueue(s, p, nullptr, &rest))) return rc;
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        return RTU_OK;
    });
}

int rtu_gather_resolve(rtu_scene *s, rtu_comm *m, const rtu_params *p, const float *d_accum, int32_t root, rtu_image *out)
{
    if (!s || !m || !p || root < 0 || root >= m->world) { rtu::set_error("rtu_gather_resolve: bad argument"); return RTU_ERR_INVALID; }
    if (m->ctx != s->ctx) { rtu::set_error("rtu_gather_resolve: the communicator belongs to another context"); return RTU_ERR_INVALID; }
    if (m->rank == root && !out) { rtu::set_error("rtu_gather_resolve: the root needs output buffers"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_gather_resolve", [&]() -> int {
        NcclApi *N = nccl_api();
        if (!N) return RTU_ERR_UNSUPPORTED;
        rtu_context *c = s->ctx;
        CU(cudaSetDevice(c->device));
        int W, H, rc;
        if ((rc = rtu_frame_dims(s, p, &W, &H))) return rc;
        if (p->spp < 1) { rtu::set_error("rtu_gather_resolve: spp must be at least 1"); return RTU_ERR_INVALID; }
        int r0 = 0, r1 = H;
        if (p->row_begin != 0 || p->row_end != 0) { r0 = p->row_begin; r1 = p->row_end; }
        if (r0 < 0 || r1 > H || r0 > r1) { rtu::set_error("rtu_gather_resolve: bad row range"); return RTU_ERR_INVALID; }
        const size_t npix = (size_t)W * H;
        const float4 *accum = d_accum ? (const float4 *)d_accum : c->fb.accum;
        if (!accum || (!d_accum && npix > c->fb.accum_n)) { rtu::set_error("rtu_gather_resolve: nothing rendered yet"); return RTU_ERR_INVALID; }
        if ((rc = rtu_ensure_image(s, npix))) return rc;
        // every rank's rows: 8 bytes per rank, all-gathered so that the root knows where each block goes
        m->h_rows[0] = r0;
        m->h_rows[1] = r1;
        CU(cudaMemcpyAsync(m->d_rows + 2 * m->rank, m->h_rows, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        NC(N->AllGather(m->d_rows + 2 * m->rank, m->d_rows, 2, ncclInt32, m->comm, c->stream));
        // the rank's own rows, finished: mean, gamma, Color24 (the other rows of the local images are never sent)
        const bool want_rgb = true, want_rgb8 = true; // the root decides what it reads; both are cheap next to the frame
        launch_resolve(c->stream, accum, (int)npix, 0.f, p->spp, want_rgb ? c->fb.d_rgb : nullptr, want_rgb8 ? c->fb.d_rgb8 : nullptr);
        if (m->rank == root) {
            CU(cudaMemcpyAsync(m->h_rows, m->d_rows, sizeof(int) * 2 * (size_t)m->world, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            std::vector<unsigned char> covered((size_t)H, 0);
            for (int r = 0; r < m->world; r++) {
                const int a = m->h_rows[2 * r], b = m->h_rows[2 * r + 1];
                if (a < 0 || b > H || a > b) { rtu::set_error("rtu_gather_resolve: a rank reported a bad row range"); return RTU_ERR_INVALID; }
                for (int y = a; y < b; y++) {
                    if (covered[y]) { rtu::set_error("rtu_gather_resolve: row ranges of the ranks overlap"); return RTU_ERR_INVALID; }
                    covered[y] = 1;
                }
            }
            for (int y = 0; y < H; y++)
                if (!covered[y]) { rtu::set_error("rtu_gather_resolve: the ranks' row ranges do not cover the image"); return RTU_ERR_INVALID; }
            NC(N->GroupStart());
            for (int r = 0; r < m->world; r++) {
                if (r == root) continue;
                const size_t a = (size_t)m->h_rows[2 * r], n = (size_t)(m->h_rows[2 * r + 1] - m->h_rows[2 * r]) * W * 3;
                if (n == 0) continue;
                NC(N->Recv(c->fb.d_rgb8 + a * W * 3, n, ncclUint8, r, m->comm, c->stream));
                NC(N->Recv(c->fb.d_rgb + a * W * 3, n, ncclFloat, r, m->comm, c->stream));
            }
            NC(N->GroupEnd());
            if (out->rgb) CU(cudaMemcpyAsync(out->rgb, c->fb.d_rgb, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            if (out->rgb8) CU(cudaMemcpyAsync(out->rgb8, c->fb.d_rgb8, npix * 3, cudaMemcpyDeviceToHost, c->stream));
            rtu_image rest = *out;
            rest.rgb = nullptr;
            rest.rgb8 = nullptr;
            if (rest.z || rest.z8 || rest.node_id || rest.face_id)
                if ((rc = rtu_resolve_enqueue(s, p, nullptr, &rest))) return rc;
            CU(cudaStreamSynchronize(c->stream));
        } else {
            const size_t n = (size_t)(r1 - r0) * W * 3;
            if (n) {
                NC(N->GroupStart());
                NC(N->Send(c->fb.d_rgb8 + (size_t)r0 * W * 3, n, ncclUint8, root, m->comm, c->stream));
                NC(N->Send(c->fb.d_rgb + (size_t)r0 * W * 3, n, ncclFloat, root, m->comm, c->stream));
                NC(N->GroupEnd());
            }
        }
        CU(cudaGetLastError());
        return RTU_OK;
    });
}

} // extern "C"
