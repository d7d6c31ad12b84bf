// rtu_render: headless replacement for the reference's main() (main.cpp:74-88) for this path.
//   rtu_render <scene.xml> [--root DIR] [--width W --height H] [--spp N] [--pattern center|ref]
//              [--mode whitted|head|photon|gather] [--bounces B] [--gi-bounces G] [--photons N] [--seed S]
//              [--out Result.png] [--zout ZBuffer.png] [--device D] [--device-bvh] [--progress]
// --device-bvh skips the host hierarchy builds (cyBVH::Build is seconds for a million triangles): the device builds an LBVH.
// --progress prints numRenderedPixels-style progress while the frame runs on the library's worker thread (BeginRender()).
// --mode head is what Render() does at the reference's HEAD (4-bounce GI + direct light; the reference uses 1024 spp);
// photon / gather are its two commented-out photon-map estimators (RenderFunctions.cpp:137-142).
// Loads the scene (LoadScene), renders it on the GPU and writes Result.png / ZBuffer.png like
// SpawnRenderThreads() does (main.cpp:59-61).  There is no window and no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rtu.h"

static int die(const char *what, int rc)
{
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, rtu_last_error());
    return 1;
}

int main(int argc, char **argv)
{
    std::string scene, root = ".", out = "Result.png", zout = "ZBuffer.png", pattern = "center", mode = "whitted";
    int width = 0, height = 0, spp = 1, bounces = 5, device = 0, gi_bounces = 4;
    long long photons = 1000000, seed = 0;
    bool device_bvh = false, progress = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--root") root = next();
        else if (a == "--width") width = atoi(next());
        else if (a == "--height") height = atoi(next());
        else if (a == "--spp") spp = atoi(next());
        else if (a == "--pattern") pattern = next();
        else if (a == "--bounces") bounces = atoi(next());
        else if (a == "--mode") mode = next();
        else if (a == "--gi-bounces") gi_bounces = atoi(next());
        else if (a == "--photons") photons = atoll(next());
        else if (a == "--seed") seed = atoll(next());
        else if (a == "--out") out = next();
        else if (a == "--zout") zout = next();
        else if (a == "--device") device = atoi(next());
        else if (a == "--device-bvh") device_bvh = true;
        else if (a == "--progress") progress = true;
        else if (a[0] != '-') scene = a;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (scene.empty()) { fprintf(stderr, "usage: rtu_render <scene.xml> [--root DIR] [--width W --height H] [--spp N] [--pattern center|ref] [--mode whitted|head|photon|gather] [--bounces B] [--gi-bounces G] [--photons N] [--seed S] [--out Result.png] [--zout ZBuffer.png]\n"); return 2; }
    rtu_host_scene *hs = nullptr;
    int rc = rtu_host_load_xml_ex(scene.c_str(), root.c_str(), device_bvh ? RTU_LOAD_DEVICE_BVH : 0u, &hs);
    if (rc) return die("rtu_host_load_xml", rc);
    if (rtu_last_error()[0]) fprintf(stderr, "%s\n", rtu_last_error());
    rtu_context *ctx = nullptr;
    if ((rc = rtu_context_create(device, nullptr, &ctx))) return die("rtu_context_create", rc);
    rtu_scene *sc = nullptr;
    if ((rc = rtu_scene_upload(ctx, rtu_host_scene_desc(hs), &sc))) return die("rtu_scene_upload", rc);
    rtu_params p;
    rtu_params_default(&p);
    p.width = width; p.height = height; p.spp = spp; p.shade_bounces = bounces; p.gi_bounces = gi_bounces; p.seed = (uint64_t)seed;
    if (mode == "whitted") p.mode = RTU_MODE_WHITTED;
    else if (mode == "head") p.mode = RTU_MODE_PATH;
    else if (mode == "photon") p.mode = RTU_MODE_PHOTON;
    else if (mode == "gather") p.mode = RTU_MODE_PHOTON_GATHER;
    else { fprintf(stderr, "unknown mode %s\n", mode.c_str()); return 2; }
    if (p.mode == RTU_MODE_PHOTON || p.mode == RTU_MODE_PHOTON_GATHER) { // GeneratePhotonMap() (main.cpp:31)
        rtu_photon_params pp;
        rtu_photon_params_default(&pp);
        pp.map_size = (uint32_t)photons;
        pp.seed = (uint64_t)seed;
        rtu_photon_stats ps;
        if ((rc = rtu_photon_map_generate(sc, &pp, &ps))) return die("rtu_photon_map_generate", rc);
        fprintf(stderr, "Photon From Light: %llu \nPhoton Scale Factor: %f \nPhoton Map Generated\n", (unsigned long long)ps.from_light, ps.scale_factor);
    }
    p.pattern = (pattern == "ref" || spp > 1) ? RTU_PATTERN_REFERENCE : RTU_PATTERN_CENTER;
    const rtu_scene_desc *d = rtu_host_scene_desc(hs);
    int W = width > 0 ? width : d->camera.width, H = height > 0 ? height : d->camera.height;
    std::vector<uint8_t> rgb8((size_t)W * H * 3), z8((size_t)W * H);
    rtu_image img;
    memset(&img, 0, sizeof img);
    img.rgb8 = rgb8.data();
    img.z8 = z8.data();
    // BeginRender(): the frame runs on the library's worker thread; this thread could draw the partial image (viewport.cpp:390-410)
    rtu_job *job = nullptr;
    auto on_progress = [](void *, int64_t done, int64_t total) { fprintf(stderr, "\r%5.1f %%", 100.0 * (double)done / (double)total); };
    if ((rc = rtu_render_async(sc, &p, &img, progress ? +on_progress : nullptr, nullptr, &job))) return die("rtu_render_async", rc);
    if ((rc = rtu_job_wait(job))) return die("rtu_render_async", rc);
    rtu_job_destroy(job);
    if (progress) fprintf(stderr, "\n");
    rtu_stats st;
    rtu_get_stats(sc, &st);
    // SaveImage / SaveZImage (main.cpp:59-61), both files encoded at the same time on worker threads
    rtu_job *w1 = nullptr, *w2 = nullptr;
    if ((rc = rtu_write_png_async(out.c_str(), rgb8.data(), W, H, 3, &w1))) return die("rtu_write_png_async", rc);
    if ((rc = rtu_write_png_async(zout.c_str(), z8.data(), W, H, 1, &w2))) return die("rtu_write_png_async", rc);
    if ((rc = rtu_job_wait(w1)) || (rc = rtu_job_wait(w2))) return die("rtu_write_png_async", rc);
    rtu_job_destroy(w1);
    rtu_job_destroy(w2);
    double rays = (double)(st.trace_rays + st.shadow_rays);
    printf("{\"width\":%d,\"height\":%d,\"spp\":%d,\"trace_rays\":%llu,\"shadow_rays\":%llu,\"device_ms\":%.3f,\"mrays_per_s\":%.2f}\n", W, H, spp,
           (unsigned long long)st.trace_rays, (unsigned long long)st.shadow_rays, st.device_ms, st.device_ms > 0 ? rays / st.device_ms * 1e-3 : 0.0);
    rtu_scene_destroy(sc);
    rtu_context_destroy(ctx);
    rtu_host_scene_destroy(hs);
    return 0;
}
