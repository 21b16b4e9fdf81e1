// Context, device-memory helpers and headless image output of the nexus_b200 C ABI.
#include "nx_common.cuh"
#include <fstream>
#include <cstdlib>

int nxi_check_overflow(nx_ctx* ctx)
{
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    NX_CUDA(ctx, cudaMemcpyAsync(ctx->hOverflow, ctx->dOverflow, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->hOverflow[1]) {      // rays with a non-finite origin or direction were answered as misses (traverse.cuh): counted, not an error
        ctx->nonfinite_rays += ctx->hOverflow[1];
        if (std::getenv("NX_DEBUG")) std::fprintf(stderr, "[nx] %u rays with non-finite origin / direction (answered as misses)\n", ctx->hOverflow[1]);
        cudaMemsetAsync(ctx->dOverflow + 1, 0, 4, ctx->stream);
    }
    if (const uint32_t dropped = *ctx->hOverflow) {
        cudaMemsetAsync(ctx->dOverflow, 0, 4, ctx->stream);
        NX_FAIL(ctx, NX_ERR_STATE, "traversal stack overflow: %u pushes beyond %u entries were refused, hits of the affected rays may be missing "
                                   "(a BVH deeper than the traversal stack)", dropped, ctx->stack_limit);
    }
    return NX_OK;
}

extern "C" {

int nx_abi_version(void) { return NX_ABI_VERSION; }

int nx_ctx_create(int device, nx_ctx** out)
{
    if (!out) return NX_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return NX_ERR_CUDA;   // no silent CPU fallback: there is none
    if (device < 0 || device >= count) return NX_ERR_INVALID;
    nx_ctx* ctx = new nx_ctx();
    ctx->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return NX_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    int prLeast = 0, prGreatest = 0;
    cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest);     // main stream: highest priority; auxiliary (shadow rays): lowest
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prGreatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->stream_aux, cudaStreamNonBlocking, prLeast) != cudaSuccess) { delete ctx; return NX_ERR_CUDA; }
    // keep freed blocks in the stream-ordered pool: builds and per-frame scratch reuse them without going to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    if (const char* t = std::getenv("NX_TRACE_TUNE")) {
        unsigned a = 0, b = 0;
        if (std::sscanf(t, "%u,%u", &a, &b) == 2) { ctx->tune_tri = ctx->tune_tri_any = a; ctx->tune_inst = ctx->tune_inst_any = b; ctx->tune_user = true; }
    }
    // L2 persistence for the top level of the scene (north_star: "L2-persistence hints for top-level nodes"): a small set-aside,
    // the window itself is set when a scene's TLAS is built (scene.cu).  NX_L2_PERSIST_MB=0 turns the hints off.
    {
        int maxPersist = 0, maxWindow = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, device);
        size_t want = 4u << 20;
        if (const char* t = std::getenv("NX_L2_PERSIST_MB")) want = (size_t)std::max(0, std::atoi(t)) << 20;
        want = std::min(want, (size_t)std::max(maxPersist, 0));
        if (want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) { ctx->l2_persist_bytes = want; ctx->l2_window_max = (size_t)std::max(maxWindow, 0); }
        cudaGetLastError();
    }
    if (const char* t = std::getenv("NX_SCENE_COLLAPSE")) {      // "mode,max_leaf_prims"
        int a = 0, b = 0;
        if (std::sscanf(t, "%d,%d", &a, &b) >= 1) { ctx->scene_collapse = a; ctx->scene_max_leaf_prims = b; }
    }
    if (const char* t = std::getenv("NX_MERGE_INSTANCES")) ctx->merge_instances = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_TRACE_GENERIC")) ctx->trace_generic = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_TLAS_REFIT")) ctx->tlas_refit = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_MERGE_MIN_PRIMS")) ctx->merge_min_prims = (uint32_t)std::max(0, std::atoi(t));
    if (const char* t = std::getenv("NX_SCENE_BLAS_SPEED")) ctx->scene_blas_speed = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_TRACE_TUNE_ANY")) {
        unsigned a = 0, b = 0;
        if (std::sscanf(t, "%u,%u", &a, &b) == 2) { ctx->tune_tri_any = a; ctx->tune_inst_any = b; }
    }
    if (const char* t = std::getenv("NX_DP_WAVES")) ctx->dp_waves = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_COLLAPSE_CTA")) ctx->collapse_cta = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_SORT")) ctx->sort_mode = std::atoi(t) != 0;
    if (const char* t = std::getenv("NX_HPLOC")) ctx->hploc_mode = std::atoi(t);
    if (const char* t = std::getenv("NX_TRACE_MODE")) ctx->trace_mode = (std::strcmp(t, "duo") == 0 || std::strcmp(t, "2") == 0) ? 2 : (std::strcmp(t, "pool") == 0 || std::strcmp(t, "1") == 0) ? 1 : 0;
    if (const char* t = std::getenv("NX_POOL_TUNE")) {
        unsigned a = 0, b = 0, c = 0, d = 0;
        if (std::sscanf(t, "%u,%u,%u,%u", &a, &b, &c, &d) == 4) { ctx->pool_node = ctx->pool_node_any = a; ctx->pool_tri = ctx->pool_tri_any = b; ctx->pool_inst = ctx->pool_inst_any = c; ctx->pool_fetch = ctx->pool_fetch_any = d; }
    }
    if (const char* t = std::getenv("NX_POOL_TUNE_ANY")) {
        unsigned a = 0, b = 0, c = 0, d = 0;
        if (std::sscanf(t, "%u,%u,%u,%u", &a, &b, &c, &d) == 4) { ctx->pool_node_any = a; ctx->pool_tri_any = b; ctx->pool_inst_any = c; ctx->pool_fetch_any = d; }
    }
    if (cudaMalloc((void**)&ctx->dOverflow, 8) != cudaSuccess || cudaMemset(ctx->dOverflow, 0, 8) != cudaSuccess || cudaMallocHost((void**)&ctx->hOverflow, 8) != cudaSuccess) {
        nx_ctx_destroy(ctx); return NX_ERR_CUDA;
    }
    ctx->hOverflow[0] = ctx->hOverflow[1] = 0;
    *out = ctx;
    return NX_OK;
}

void nx_ctx_destroy(nx_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream_aux);
    cudaFree(ctx->dOverflow); cudaFreeHost(ctx->hOverflow);
    for (int k = 0; k < ctx->buildStreamCount; k++) { cudaStreamSynchronize(ctx->buildStreams[k]); cudaStreamDestroy(ctx->buildStreams[k]); }
    if (ctx->stagePinned) cudaFreeHost(ctx->stagePinned);
    for (nx_bump& b : ctx->buildWsStore) cudaFree(b.base);
    cudaFree(ctx->poolSpill[0]); cudaFree(ctx->poolSpill[1]);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream_aux);
    // the blocks the stream-ordered pool kept for reuse (release threshold raised in nx_ctx_create) go back to the driver; blocks another
    // live context of this device still holds are not touched by a trim
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(pool, 0); }
    delete ctx;
}

const char* nx_last_error(const nx_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
int nx_ctx_sm_count(const nx_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
void* nx_ctx_stream(nx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int nx_ctx_set_trace_tuning(nx_ctx* ctx, uint32_t tri_lanes, uint32_t inst_lanes)
{
    if (!ctx || tri_lanes > 32 || inst_lanes > 32) return NX_ERR_INVALID;
    ctx->tune_tri = tri_lanes; ctx->tune_inst = inst_lanes; ctx->tune_user = true;
    if (!std::getenv("NX_TRACE_TUNE_ANY")) { ctx->tune_tri_any = tri_lanes; ctx->tune_inst_any = inst_lanes; }
    return NX_OK;
}

int nx_ctx_set_scene_collapse(nx_ctx* ctx, int collapse, int max_leaf_prims)
{
    if (!ctx || (collapse != NX_COLLAPSE_REFERENCE_GPU && collapse != NX_COLLAPSE_SAH_OPTIMAL) || max_leaf_prims < 0 || max_leaf_prims > 3) return NX_ERR_INVALID;
    ctx->scene_collapse = collapse; ctx->scene_max_leaf_prims = max_leaf_prims;
    return NX_OK;
}

int nx_ctx_set_tlas_refit(nx_ctx* ctx, int enabled)
{
    if (!ctx) return NX_ERR_INVALID;
    ctx->tlas_refit = enabled ? 1 : 0;
    return NX_OK;
}

int nx_ctx_set_trace_mode(nx_ctx* ctx, int mode)
{
    if (!ctx || (mode != NX_TRACE_LANE && mode != NX_TRACE_POOL && mode != NX_TRACE_DUO && mode != NX_TRACE_LANE_GENERAL)) return NX_ERR_INVALID;
    ctx->trace_mode = mode == NX_TRACE_LANE_GENERAL ? NX_TRACE_LANE : mode;
    ctx->trace_generic = mode == NX_TRACE_LANE_GENERAL ? 1 : 0;
    return NX_OK;
}

int nx_ctx_set_pool_tuning(nx_ctx* ctx, int any_hit, uint32_t node_rays, uint32_t tri_rays, uint32_t inst_rays, uint32_t fetch_rays)
{
    if (!ctx || !node_rays || !tri_rays || !inst_rays || !fetch_rays) return NX_ERR_INVALID;
    if (any_hit) { ctx->pool_node_any = node_rays; ctx->pool_tri_any = tri_rays; ctx->pool_inst_any = inst_rays; ctx->pool_fetch_any = fetch_rays; }
    else { ctx->pool_node = node_rays; ctx->pool_tri = tri_rays; ctx->pool_inst = inst_rays; ctx->pool_fetch = fetch_rays; }
    return NX_OK;
}

int nx_ctx_set_stack_limit(nx_ctx* ctx, uint32_t entries)
{
    if (!ctx || entries < 2 || entries > 40) return NX_ERR_INVALID;
    ctx->stack_limit = entries;
    return NX_OK;
}

int nx_ctx_set_instance_merging(nx_ctx* ctx, int enabled)
{
    if (!ctx) return NX_ERR_INVALID;
    ctx->merge_instances = enabled ? 1 : 0;
    return NX_OK;
}

int nx_ctx_set_sphere_cull(nx_ctx* ctx, int enabled)
{
    if (!ctx) return NX_ERR_INVALID;
    ctx->tune_sphere = enabled ? 1u : 0u;
    return NX_OK;
}

int nx_ctx_synchronize(nx_ctx* ctx)
{
    if (!ctx) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    for (int k = 0; k < ctx->buildStreamCount; k++) NX_CUDA(ctx, cudaStreamSynchronize(ctx->buildStreams[k]));
    NX_CUDA(ctx, cudaGetLastError());
    return nxi_check_overflow(ctx);
}

int nx_malloc(nx_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return NX_OK;
}
int nx_free(nx_ctx* ctx, void* dev)
{
    if (!ctx) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    NX_CUDA(ctx, cudaFree(dev));
    return NX_OK;
}
int nx_memcpy_h2d(nx_ctx* ctx, void* dev, const void* host, size_t bytes)
{
    if (!ctx) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NX_OK;
}
int nx_memcpy_d2h(nx_ctx* ctx, void* host, const void* dev, size_t bytes)
{
    if (!ctx) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NX_OK;
}

// ------------------------------------------------------------------------------------------ image output ----
// Both writers take the renderer's own layout: row 0 of `rgb` is the BOTTOM row of the image (generate_kernel maps pixel row 0 to the
// camera's lower-left corner like the reference, PathTracer.cu:74-83), i.e. exactly what nx_renderer_read_accum returns.
// PFM: "PF\nW H\n-1.0\n" then the rows bottom to top, little-endian float RGB - the input order.
int nx_write_pfm(const char* path, const float* rgb, uint32_t w, uint32_t h)
{
    if (!path || !rgb) return NX_ERR_INVALID;
    std::ofstream f(path, std::ios::binary);
    if (!f) return NX_ERR_INVALID;
    f << "PF\n" << w << " " << h << "\n-1.0\n";
    f.write((const char*)rgb, 12 * (size_t)w * h);
    return f ? NX_OK : NX_ERR_INVALID;
}

// Minimal OpenEXR 2.0 writer: single-part scanline image, three FLOAT channels (B, G, R in file order), no compression.
// EXR scanline 0 is the TOP of the image, so scanline y is input row h - 1 - y.
int nx_write_exr(const char* path, const float* rgb, uint32_t w, uint32_t h)
{
    if (!path || !rgb || !w || !h) return NX_ERR_INVALID;
    std::vector<uint8_t> hd;
    auto put = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; hd.insert(hd.end(), b, b + n); };
    auto str = [&](const char* s) { put(s, std::strlen(s) + 1); };
    auto i32 = [&](int32_t v) { put(&v, 4); };
    auto f32 = [&](float v) { put(&v, 4); };
    i32(20000630); i32(2);
    str("channels"); str("chlist"); i32(3 * 18 + 1);
    for (const char* c : {"B", "G", "R"}) { str(c); i32(2 /* FLOAT */); uint8_t lin[4] = {0, 0, 0, 0}; put(lin, 4); i32(1); i32(1); }
    hd.push_back(0);
    str("compression"); str("compression"); i32(1); hd.push_back(0);
    str("dataWindow"); str("box2i"); i32(16); i32(0); i32(0); i32((int32_t)w - 1); i32((int32_t)h - 1);
    str("displayWindow"); str("box2i"); i32(16); i32(0); i32(0); i32((int32_t)w - 1); i32((int32_t)h - 1);
    str("lineOrder"); str("lineOrder"); i32(1); hd.push_back(0);
    str("pixelAspectRatio"); str("float"); i32(4); f32(1.0f);
    str("screenWindowCenter"); str("v2f"); i32(8); f32(0.0f); f32(0.0f);
    str("screenWindowWidth"); str("float"); i32(4); f32(1.0f);
    hd.push_back(0);
    std::ofstream f(path, std::ios::binary);
    if (!f) return NX_ERR_INVALID;
    f.write((const char*)hd.data(), hd.size());
    const uint64_t rowBytes = 12ull * w, chunk = 8 + rowBytes;
    uint64_t off = hd.size() + 8ull * h;
    for (uint32_t y = 0; y < h; y++) { f.write((const char*)&off, 8); off += chunk; }
    std::vector<float> row(3 * (size_t)w);
    for (uint32_t y = 0; y < h; y++) {
        const float* src = rgb + 3 * (size_t)(h - 1 - y) * w;
        for (uint32_t x = 0; x < w; x++) { row[x] = src[3 * x + 2]; row[w + x] = src[3 * x + 1]; row[2 * (size_t)w + x] = src[3 * x]; }
        int32_t yy = (int32_t)y, sz = (int32_t)rowBytes;
        f.write((const char*)&yy, 4); f.write((const char*)&sz, 4); f.write((const char*)row.data(), rowBytes);
    }
    return f ? NX_OK : NX_ERR_INVALID;
}

} // extern "C"
