// Wavefront path tracer for sm_100a: generate -> trace -> shade (logic + OpenPBR material + NEE/MIS) -> {trace, trace shadow}
// -> accumulate.  Replaces the six kernels of src/Cuda/PathTracer/PathTracer.cu:60-549 and PathTracer::Render
// (src/Renderer/PathTracer.cpp:166-200) of the reference.
//
// Differences that matter for throughput:
//   * Logic and Material are one kernel: the miss / Russian-roulette decisions are taken in registers right before shading
//     instead of through a 36 B/ray material queue round trip;
//   * path throughput and last pdf travel with the ray (one 16-byte record in queue order) instead of being gathered and
//     scattered through per-pixel arrays;
//   * queue appends are one atomic per warp (ballot + popc prefix) at a convergent point;
//   * radiance goes straight into the float sum buffer with RED.ADD (no per-frame radiance buffer, no accumulate pass over
//     W*H pixels per frame); the mean is taken on read-out, which is also what the multi-GPU reduction wants;
//   * all kernels are persistent and sized from the SM count; queue lengths are read on the device, so no kernel is
//     launched over W*H threads for a nearly empty queue and the host never waits inside a frame;
//   * rays are 32-byte records, hits 20-byte records: 16-byte vector loads in queue order.
#include "wave.cuh"   // shared wavefront state; generation / shading / display kernels live in shade.cu (fast-math unit)
#include "traverse_pool.cuh"
#include "traverse_duo.cuh"
#include <algorithm>
namespace {

// ------------------------------------------------------------------------------------------ trace kernels ----
// Persistent warps over a device-side queue; the loop itself is trace_loop() in traverse.cuh.
struct ClosestSink {
    nx_hit* hits;
    __device__ __forceinline__ void finish(const TraceScene& sc, uint32_t rayIdx, uint32_t, float t, float u, float v, uint32_t prim, uint32_t slot, bool)
    {
        nx_hit h; h.t = t; h.u = u; h.v = v; h.prim = prim;
        h.instance = hit_instance(sc, slot);
        hits[rayIdx] = h;
    }
};
// Any-hit.  occluded != null: write occlusion flags (parity hook); else add the queued radiance to the pixel when unoccluded
// (TraceShadowKernel's fused accumulate, PathTracer.cu:115-122 / BVH8Traversal.cuh:517-519).
struct AnySink {
    uint8_t* occluded; const float4* radiance; float* accum;
    __device__ __forceinline__ void finish(const TraceScene&, uint32_t rayIdx, uint32_t pixel, float, float, float, uint32_t, uint32_t, bool occ)
    {
        if (occluded) occluded[rayIdx] = occ ? 1 : 0;
        else if (!occ) { const float4 L = __ldg(radiance + rayIdx); add_radiance(accum, pixel, f3(L.x, L.y, L.z)); }
    }
};

#ifndef NX_TRACE_MIN_BLOCKS_DIRECT
#define NX_TRACE_MIN_BLOCKS_DIRECT NX_TRACE_MIN_BLOCKS
#endif
template <bool STATS, int KIND>
__global__ void __launch_bounds__(NX_TRACE_BLOCK, KIND == NX_SCENE_DIRECT ? NX_TRACE_MIN_BLOCKS_DIRECT : NX_TRACE_MIN_BLOCKS) trace_closest_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                         uint32_t* cursor, nx_hit* __restrict__ hits, TraceStats* stats, TraceTuning tune)
{
    __shared__ __align__(16) uint32_t smem[(KIND == NX_SCENE_DIRECT ? NX_TRACE_SMEM_BYTES_DIRECT : NX_TRACE_SMEM_BYTES) / 4];
    ClosestSink sink{hits};
    trace_loop<false, STATS, KIND>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, smem, sink, stats);
}

template <bool STATS, int KIND>
__global__ void __launch_bounds__(NX_TRACE_BLOCK, KIND == NX_SCENE_DIRECT ? NX_TRACE_MIN_BLOCKS_DIRECT : NX_TRACE_MIN_BLOCKS) trace_any_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                     uint32_t* cursor, uint8_t* occluded, const float4* __restrict__ radiance, float* accum,
                                                                     TraceStats* stats, TraceTuning tune)
{
    __shared__ __align__(16) uint32_t smem[(KIND == NX_SCENE_DIRECT ? NX_TRACE_SMEM_BYTES_DIRECT : NX_TRACE_SMEM_BYTES) / 4];
    AnySink sink{occluded, radiance, accum};
    trace_loop<true, STATS, KIND>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, smem, sink, stats);
}

// Two rays per lane (traverse_duo.cuh).
template <bool STATS>
__global__ void __launch_bounds__(NX_DUO_BLOCK, NX_DUO_MIN_BLOCKS) trace_closest_duo_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                         uint32_t* cursor, nx_hit* __restrict__ hits, TraceStats* stats, TraceTuning tune)
{
    __shared__ __align__(16) uint32_t smem[NX_DUO_SMEM_BYTES / 4];
    ClosestSink sink{hits};
    trace_loop_duo<false, STATS>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, smem, sink, stats);
}
template <bool STATS>
__global__ void __launch_bounds__(NX_DUO_BLOCK, NX_DUO_MIN_BLOCKS) trace_any_duo_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                     uint32_t* cursor, uint8_t* occluded, const float4* __restrict__ radiance, float* accum,
                                                                     TraceStats* stats, TraceTuning tune)
{
    __shared__ __align__(16) uint32_t smem[NX_DUO_SMEM_BYTES / 4];
    AnySink sink{occluded, radiance, accum};
    trace_loop_duo<true, STATS>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, smem, sink, stats);
}

// Ray-pool versions (traverse_pool.cuh): 64 rays per warp in dynamic shared memory, spill stacks of this launch's warps in `spill`.
template <bool STATS>
__global__ void __launch_bounds__(NX_POOL_BLOCK, NX_POOL_MIN_BLOCKS) trace_closest_pool_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                         uint32_t* cursor, nx_hit* __restrict__ hits, TraceStats* stats, PoolTuning tune, uint2* spill)
{
    extern __shared__ __align__(16) unsigned char pool_smem[];
    PoolWarp& pw = reinterpret_cast<PoolWarp*>(pool_smem)[threadIdx.x >> 5];
    ClosestSink sink{hits};
    trace_pool_loop<false, STATS>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, pw, sink, stats,
                                  spill + (size_t)(blockIdx.x * NX_POOL_WARPS + (threadIdx.x >> 5)) * (NX_POOL_R * NX_POOL_SPILL));
}

template <bool STATS>
__global__ void __launch_bounds__(NX_POOL_BLOCK, NX_POOL_MIN_BLOCKS) trace_any_pool_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                     uint32_t* cursor, uint8_t* occluded, const float4* __restrict__ radiance, float* accum,
                                                                     TraceStats* stats, PoolTuning tune, uint2* spill)
{
    extern __shared__ __align__(16) unsigned char pool_smem[];
    PoolWarp& pw = reinterpret_cast<PoolWarp*>(pool_smem)[threadIdx.x >> 5];
    AnySink sink{occluded, radiance, accum};
    trace_pool_loop<true, STATS>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, pw, sink, stats,
                                 spill + (size_t)(blockIdx.x * NX_POOL_WARPS + (threadIdx.x >> 5)) * (NX_POOL_R * NX_POOL_SPILL));
}

__global__ void frame_totals_kernel(WaveBuffers wb, uint32_t pathLength)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long e = 0, s = 0, h = 0;
    for (uint32_t b = 0; b <= pathLength && b < kMaxBounce; b++) { e += wb.counters->extCount[b]; s += wb.counters->shCount[b]; h += wb.counters->shaded[b]; }
    wb.totals->ext += e; wb.totals->shadow += s; wb.totals->shaded += h; wb.totals->frames += 1;
}

// Pixel query (LogicKernel / MaterialKernel at bounce 1, PathTracer.cu:150-151, 459-460): the primary hit of one pixel.  The
// primary-ray queue is in generation order, so the pixel's queue slot is known on the host (pixel_to_slot, wave.cuh).
__global__ void pixel_query_kernel(const nx_hit* __restrict__ hits, uint32_t slot, int32_t* out)
{
    const nx_hit h = hits[slot];
    *out = h.prim == NX_INVALID ? -1 : (int32_t)h.instance;
}

TraceTuning trace_tuning(const nx_ctx* ctx, bool any = false)
{
    TraceTuning t; t.triLanes = any ? ctx->tune_tri_any : ctx->tune_tri; t.instLanes = any ? ctx->tune_inst_any : ctx->tune_inst; t.sphereCull = ctx->tune_sphere; t.k47 = 0x47000000u;
    t.stackLimit = std::min<uint32_t>(std::max<uint32_t>(ctx->stack_limit, NX_STACK_SHARED), NX_STACK_TOTAL);
    return t;
}
PoolTuning pool_tuning(const nx_ctx* ctx, bool any = false)
{
    PoolTuning t;
    t.nodeLanes = any ? ctx->pool_node_any : ctx->pool_node; t.triLanes = any ? ctx->pool_tri_any : ctx->pool_tri;
    t.instLanes = any ? ctx->pool_inst_any : ctx->pool_inst; t.fetchLanes = any ? ctx->pool_fetch_any : ctx->pool_fetch;
    t.sphereCull = ctx->tune_sphere; t.k47 = 0x47000000u; t.stackLimit = std::min<uint32_t>(ctx->stack_limit, NX_STACK_TOTAL);
    return t;
}

// Persistent grid of a kernel on this context's device: resident CTAs per SM x SM count, cached in the context.
int persistent_grid(nx_ctx* ctx, const void* fn, int block, size_t smem, int slot)
{
    int& cache = ctx->gridCache[slot];
    if (cache) return cache;
    if (smem > 48u * 1024u) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, block, smem);
    if (perSm < 1) perSm = 1;
    cache = perSm * ctx->sm_count;
    return cache;
}

// global spill stacks of the ray-pool kernels: one region per trace stream (closest-hit and shadow traces overlap)
int pool_spill(nx_ctx* ctx, int which, int grid, uint2** out)
{
    const size_t warps = (size_t)grid * NX_POOL_WARPS;
    if (ctx->poolSpillWarps[which] < warps) {
        cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream_aux);
        cudaFree(ctx->poolSpill[which]); ctx->poolSpill[which] = nullptr; ctx->poolSpillWarps[which] = 0;
        NX_CUDA(ctx, cudaMalloc(&ctx->poolSpill[which], warps * NX_POOL_R * NX_POOL_SPILL * sizeof(uint2)));
        ctx->poolSpillWarps[which] = warps;
    }
    *out = (uint2*)ctx->poolSpill[which];
    return NX_OK;
}

inline int scene_kind(const TraceScene& sc) { return sc.mergedSlot == NX_INVALID ? NX_SCENE_TWO_LEVEL : (sc.direct ? NX_SCENE_DIRECT : NX_SCENE_MIXED); }

// One closest-hit launch over a ray queue (count immediate or read on the device) in the context's traversal mode.
template <bool STATS>
int launch_closest(nx_ctx* ctx, cudaStream_t st, const TraceScene& sc, const nx_ray* q, uint32_t nImm, const uint32_t* nPtr, uint32_t* cursor, nx_hit* hits, TraceStats* stats)
{
    if (ctx->trace_mode == 1) {
        const int grid = persistent_grid(ctx, (const void*)trace_closest_pool_kernel<STATS>, NX_POOL_BLOCK, NX_POOL_SMEM_BYTES, STATS ? 5 : 4);
        uint2* spill = nullptr; int rc = pool_spill(ctx, 0, grid, &spill); if (rc) return rc;
        trace_closest_pool_kernel<STATS><<<grid, NX_POOL_BLOCK, NX_POOL_SMEM_BYTES, st>>>(sc, q, nImm, nPtr, cursor, hits, stats, pool_tuning(ctx), spill);
    } else if (ctx->trace_mode == 2) {
        const int grid = persistent_grid(ctx, (const void*)trace_closest_duo_kernel<STATS>, NX_DUO_BLOCK, 0, STATS ? 9 : 8);
        TraceTuning t = trace_tuning(ctx); t.stackLimit = std::min<uint32_t>(std::max<uint32_t>(ctx->stack_limit, NX_DUO_STACK), NX_STACK_TOTAL);
        trace_closest_duo_kernel<STATS><<<grid, NX_DUO_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, hits, stats, t);
    } else {
        // the loop specialised for what the scene holds (traverse.cuh); NX_TRACE_GENERIC=1 forces the general one (tests: same bytes)
        const int kind = ctx->trace_generic ? NX_SCENE_MIXED : scene_kind(sc);
        TraceTuning t = trace_tuning(ctx);
        if (kind == NX_SCENE_DIRECT && !ctx->tune_user) { t.triLanes = 8; t.instLanes = 4; }
        if (kind == NX_SCENE_DIRECT) {
            const int grid = persistent_grid(ctx, (const void*)trace_closest_kernel<STATS, NX_SCENE_DIRECT>, NX_TRACE_BLOCK, 0, STATS ? 17 : 16);
            trace_closest_kernel<STATS, NX_SCENE_DIRECT><<<grid, NX_TRACE_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, hits, stats, t);
        } else if (kind == NX_SCENE_TWO_LEVEL) {
            const int grid = persistent_grid(ctx, (const void*)trace_closest_kernel<STATS, NX_SCENE_TWO_LEVEL>, NX_TRACE_BLOCK, 0, STATS ? 19 : 18);
            trace_closest_kernel<STATS, NX_SCENE_TWO_LEVEL><<<grid, NX_TRACE_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, hits, stats, t);
        } else {
            const int grid = persistent_grid(ctx, (const void*)trace_closest_kernel<STATS, NX_SCENE_MIXED>, NX_TRACE_BLOCK, 0, STATS ? 1 : 0);
            trace_closest_kernel<STATS, NX_SCENE_MIXED><<<grid, NX_TRACE_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, hits, stats, t);
        }
    }
    return NX_OK;
}
template <bool STATS>
int launch_any(nx_ctx* ctx, cudaStream_t st, const TraceScene& sc, const nx_ray* q, uint32_t nImm, const uint32_t* nPtr, uint32_t* cursor,
               uint8_t* occ, const float4* rad, float* accum, TraceStats* stats)
{
    if (ctx->trace_mode == 1) {
        const int grid = persistent_grid(ctx, (const void*)trace_any_pool_kernel<STATS>, NX_POOL_BLOCK, NX_POOL_SMEM_BYTES, STATS ? 7 : 6);
        uint2* spill = nullptr; int rc = pool_spill(ctx, 1, grid, &spill); if (rc) return rc;
        trace_any_pool_kernel<STATS><<<grid, NX_POOL_BLOCK, NX_POOL_SMEM_BYTES, st>>>(sc, q, nImm, nPtr, cursor, occ, rad, accum, stats, pool_tuning(ctx, true), spill);
    } else if (ctx->trace_mode == 2) {
        const int grid = persistent_grid(ctx, (const void*)trace_any_duo_kernel<STATS>, NX_DUO_BLOCK, 0, STATS ? 11 : 10);
        TraceTuning t = trace_tuning(ctx, true); t.stackLimit = std::min<uint32_t>(std::max<uint32_t>(ctx->stack_limit, NX_DUO_STACK), NX_STACK_TOTAL);
        trace_any_duo_kernel<STATS><<<grid, NX_DUO_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, occ, rad, accum, stats, t);
    } else {
        const int kind = ctx->trace_generic ? NX_SCENE_MIXED : scene_kind(sc);
        TraceTuning t = trace_tuning(ctx, true);
        if (kind == NX_SCENE_DIRECT && !ctx->tune_user) { t.triLanes = 8; t.instLanes = 4; }
        if (kind == NX_SCENE_DIRECT) {
            const int grid = persistent_grid(ctx, (const void*)trace_any_kernel<STATS, NX_SCENE_DIRECT>, NX_TRACE_BLOCK, 0, STATS ? 21 : 20);
            trace_any_kernel<STATS, NX_SCENE_DIRECT><<<grid, NX_TRACE_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, occ, rad, accum, stats, t);
        } else if (kind == NX_SCENE_TWO_LEVEL) {
            const int grid = persistent_grid(ctx, (const void*)trace_any_kernel<STATS, NX_SCENE_TWO_LEVEL>, NX_TRACE_BLOCK, 0, STATS ? 23 : 22);
            trace_any_kernel<STATS, NX_SCENE_TWO_LEVEL><<<grid, NX_TRACE_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, occ, rad, accum, stats, t);
        } else {
            const int grid = persistent_grid(ctx, (const void*)trace_any_kernel<STATS, NX_SCENE_MIXED>, NX_TRACE_BLOCK, 0, STATS ? 3 : 2);
            trace_any_kernel<STATS, NX_SCENE_MIXED><<<grid, NX_TRACE_BLOCK, 0, st>>>(sc, q, nImm, nPtr, cursor, occ, rad, accum, stats, t);
        }
    }
    return NX_OK;
}

} // namespace

struct nx_renderer {
    nx_ctx* ctx = nullptr;
    uint32_t width = 0, height = 0;
    WaveBuffers wb{};
    uint32_t frames = 0;            // frames in the accumulation
    cudaEvent_t evStart = nullptr, evStop = nullptr, evShade = nullptr, evShadow[2] = {nullptr, nullptr};
    bool timed = false;
    uint32_t launches = 0;
    uint32_t pathLengthLast = 0;
    // optional per-kernel profiling (nx_renderer_set_profiling): event pairs per launch, grouped by kernel
    int profFlags = 0;
    std::vector<cudaEvent_t> evPool; size_t evUsed = 0;
    std::vector<std::pair<int, size_t>> evLaunches;   // (kernel class, index of the start event; stop = index + 1)
    TraceStats* dWork = nullptr;                      // [0] closest-hit traversal work, [1] any-hit traversal work
    // pipelined display read-back (nx_renderer_present): two device RGBA8 images, a copy stream, pinned totals per ticket
    cudaStream_t copyStream = nullptr;
    uint32_t* dRgba[2] = {nullptr, nullptr};
    cudaEvent_t evResolved[2] = {nullptr, nullptr}, evCopied[2] = {nullptr, nullptr};
    bool slotUsed[2] = {false, false};
    WaveTotals* hTotals = nullptr;                    // pinned, [2]
    int presentNext = 0;
    // pixel query (SetPixelQuery / SynchronizePixelQuery)
    int32_t* dQuery = nullptr; int32_t* hQuery = nullptr;   // device slot, pinned host mirror
    int64_t queryPixel = -1; bool queryPending = false;

    cudaEvent_t next_event() { if (evUsed == evPool.size()) { cudaEvent_t e; cudaEventCreate(&e); evPool.push_back(e); } return evPool[evUsed++]; }
    void prof_begin(int cls, cudaStream_t st) { if (profFlags & 1) { evLaunches.push_back({cls, evUsed}); cudaEventRecord(next_event(), st); next_event(); } }
    void prof_end(cudaStream_t st) { if (profFlags & 1) cudaEventRecord(evPool[evLaunches.back().second + 1], st); }
};

int nxi_trace_closest(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, nx_hit* dHits, float* outMs)
{
    DeviceGuard guard(ctx->device);
    uint32_t* cursor = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&cursor, 4, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (outMs) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream); }
    int rc = launch_closest<false>(ctx, ctx->stream, sc, dRays, n, nullptr, cursor, dHits, nullptr); if (rc) return rc;
    if (outMs) { cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1); cudaEventElapsedTime(outMs, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    cudaFreeAsync(cursor, ctx->stream);
    NX_CUDA(ctx, cudaGetLastError());
    return nxi_check_overflow(ctx);
}

int nxi_trace_any(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, uint8_t* dOcc, float* outMs)
{
    DeviceGuard guard(ctx->device);
    uint32_t* cursor = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&cursor, 4, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (outMs) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream); }
    int rc = launch_any<false>(ctx, ctx->stream, sc, dRays, n, nullptr, cursor, dOcc, nullptr, nullptr, nullptr); if (rc) return rc;
    if (outMs) { cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1); cudaEventElapsedTime(outMs, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    cudaFreeAsync(cursor, ctx->stream);
    NX_CUDA(ctx, cudaGetLastError());
    return nxi_check_overflow(ctx);
}

// Traversal work counters for the roofline's algorithmic-byte figure: nodes, triangles, instances per ray.
extern "C" int nx_trace_stats(nx_scene* s, const nx_ray* dRays, uint32_t n, nx_hit* dHits, uint64_t out4[4])
{
    if (!s || !dRays || !dHits || !out4) return NX_ERR_INVALID;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    DSceneView v; int rc = nxi_scene_view(s, &v); if (rc) return rc;
    uint32_t* cursor = nullptr; TraceStats* st = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&cursor, 4, ctx->stream)); NX_CUDA(ctx, cudaMallocAsync((void**)&st, sizeof(TraceStats), ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream)); NX_CUDA(ctx, cudaMemsetAsync(st, 0, sizeof(TraceStats), ctx->stream));
    rc = launch_closest<true>(ctx, ctx->stream, v.trace, dRays, n, nullptr, cursor, dHits, st); if (rc) return rc;
    TraceStats h{};
    NX_CUDA(ctx, cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out4[0] = h.nodes; out4[1] = h.tris; out4[2] = h.insts; out4[3] = h.rays;
    cudaFreeAsync(cursor, ctx->stream); cudaFreeAsync(st, ctx->stream);
    return NX_OK;
}

namespace {

int free_buffers(nx_renderer* r)
{
    nx_ctx* ctx = r->ctx;
    cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream_aux);
    WaveBuffers& w = r->wb;
    cudaFree(w.ext[0]); cudaFree(w.ext[1]); cudaFree(w.state[0]); cudaFree(w.state[1]); cudaFree(w.hits); cudaFree(w.shadow[0]); cudaFree(w.shadow[1]);
    cudaFree(w.shadowRad[0]); cudaFree(w.shadowRad[1]); cudaFree(w.accum); cudaFree(w.counters); cudaFree(w.totals);
    w = WaveBuffers{};
    if (r->copyStream) cudaStreamSynchronize(r->copyStream);
    for (int k = 0; k < 2; k++) { cudaFree(r->dRgba[k]); r->dRgba[k] = nullptr; r->slotUsed[k] = false; }
    r->queryPixel = -1; r->queryPending = false;
    return NX_OK;
}

// PathTracer::Reset (src/Renderer/PathTracer.cpp:61-159): 32+32+16+16+20+32+16 = 164 B of queues per pixel + 12 B accumulation
int alloc_buffers(nx_renderer* r, uint32_t w, uint32_t h)
{
    nx_ctx* ctx = r->ctx;
    const size_t px = (size_t)w * h;
    WaveBuffers& b = r->wb;
    NX_CUDA(ctx, cudaMalloc((void**)&b.ext[0], sizeof(nx_ray) * px)); NX_CUDA(ctx, cudaMalloc((void**)&b.ext[1], sizeof(nx_ray) * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.state[0], 16 * px)); NX_CUDA(ctx, cudaMalloc((void**)&b.state[1], 16 * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.hits, sizeof(nx_hit) * px));
    for (int k = 0; k < 2; k++) { NX_CUDA(ctx, cudaMalloc((void**)&b.shadow[k], sizeof(nx_ray) * px)); NX_CUDA(ctx, cudaMalloc((void**)&b.shadowRad[k], 16 * px)); }
    NX_CUDA(ctx, cudaMalloc((void**)&b.accum, 12 * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.counters, sizeof(WaveCounters))); NX_CUDA(ctx, cudaMalloc((void**)&b.totals, sizeof(WaveTotals)));
    NX_CUDA(ctx, cudaMemsetAsync(b.accum, 0, 12 * px, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(b.totals, 0, sizeof(WaveTotals), ctx->stream));
    r->width = w; r->height = h; r->frames = 0;
    return NX_OK;
}

} // namespace

extern "C" {

int nx_renderer_create(nx_ctx* ctx, uint32_t width, uint32_t height, nx_renderer** out)
{
    if (!ctx || !out || !width || !height) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    nx_renderer* r = new nx_renderer(); r->ctx = ctx;
    int rc = alloc_buffers(r, width, height);
    if (rc) { delete r; return rc; }
    cudaEventCreate(&r->evStart); cudaEventCreate(&r->evStop);
    cudaEventCreateWithFlags(&r->evShade, cudaEventDisableTiming);
    for (int k = 0; k < 2; k++) cudaEventCreateWithFlags(&r->evShadow[k], cudaEventDisableTiming);
    if (cudaMalloc((void**)&r->dWork, 2 * sizeof(TraceStats)) != cudaSuccess || cudaMemset(r->dWork, 0, 2 * sizeof(TraceStats)) != cudaSuccess) {
        ctx->error = "nx_renderer_create: cudaMalloc failed"; nx_renderer_destroy(r); return NX_ERR_CUDA;
    }
    bool ok = cudaStreamCreateWithFlags(&r->copyStream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < 2 && ok; k++)
        ok = cudaEventCreateWithFlags(&r->evResolved[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&r->evCopied[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&r->hTotals, 2 * sizeof(WaveTotals)) == cudaSuccess && cudaMallocHost((void**)&r->hQuery, 4) == cudaSuccess
            && cudaMalloc((void**)&r->dQuery, 4) == cudaSuccess && cudaMemset(r->dQuery, 0xff, 4) == cudaSuccess;
    if (!ok) { ctx->error = "nx_renderer_create: read-back resources"; nx_renderer_destroy(r); return NX_ERR_CUDA; }
    *r->hQuery = -1;
    *out = r;
    return NX_OK;
}

void nx_renderer_destroy(nx_renderer* r)
{
    if (!r) return;
    DeviceGuard guard(r->ctx->device);
    free_buffers(r);
    cudaEventDestroy(r->evStart); cudaEventDestroy(r->evStop); cudaEventDestroy(r->evShade); cudaEventDestroy(r->evShadow[0]); cudaEventDestroy(r->evShadow[1]);
    for (cudaEvent_t e : r->evPool) cudaEventDestroy(e);
    cudaFree(r->dWork);
    for (int k = 0; k < 2; k++) { if (r->evResolved[k]) cudaEventDestroy(r->evResolved[k]); if (r->evCopied[k]) cudaEventDestroy(r->evCopied[k]); }
    if (r->copyStream) cudaStreamDestroy(r->copyStream);
    cudaFreeHost(r->hTotals); cudaFreeHost(r->hQuery); cudaFree(r->dQuery);
    delete r;
}

int nx_renderer_resize(nx_renderer* r, uint32_t width, uint32_t height)
{
    if (!r || !width || !height) return NX_ERR_INVALID;
    if (width == r->width && height == r->height) return NX_OK;
    DeviceGuard guard(r->ctx->device);
    free_buffers(r);
    r->width = r->height = 0;                       // until the new buffers exist every call on this renderer fails with NX_ERR_STATE
    const int rc = alloc_buffers(r, width, height);
    if (rc) { free_buffers(r); r->width = r->height = 0; }
    return rc;
}

int nx_renderer_reset_accumulation(nx_renderer* r)
{
    if (!r) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMemsetAsync(r->wb.accum, 0, 12 * (size_t)r->width * r->height, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(r->wb.totals, 0, sizeof(WaveTotals), ctx->stream));
    r->frames = 0; r->timed = false;
    return NX_OK;
}

int nx_renderer_render(nx_renderer* r, nx_scene* scene, uint32_t firstFrame, uint32_t nFrames)
{
    if (!r || !scene || scene->ctx != r->ctx) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    if (!r->width || !r->wb.accum) NX_FAIL(ctx, NX_ERR_STATE, "the renderer has no buffers (a resize failed)");
    if (scene->width != r->width || scene->height != r->height) NX_FAIL(ctx, NX_ERR_INVALID, "scene resolution %ux%u != renderer %ux%u", scene->width, scene->height, r->width, r->height);
    DeviceGuard guard(ctx->device);
    DSceneView sv; int rc = nxi_scene_view(scene, &sv); if (rc) return rc;
    // with per-launch profiling events on, the shadow trace runs on the main stream too: kernels then execute one at a time and
    // the event pairs measure each kernel alone (comparable with ncu's serialised launch list) instead of two overlapping ones
    cudaStream_t s = ctx->stream, sa = (r->profFlags & 1) ? ctx->stream : ctx->stream_aux;
    const uint32_t L = sv.pathLength;
    const bool work = (r->profFlags & 2) != 0;
    const int gShade = nxi_shade_grid(ctx);
    const int gGen = ctx->sm_count * 8;
    WaveBuffers& wb = r->wb;
    r->evUsed = 0; r->evLaunches.clear();
    if (work) NX_CUDA(ctx, cudaMemsetAsync(r->dWork, 0, 2 * sizeof(TraceStats), s));

    auto closest = [&](const nx_ray* q, uint32_t b) {
        r->prof_begin(1, s);
        if (work) launch_closest<true>(ctx, s, sv.trace, q, 0, &wb.counters->extCount[b], &wb.counters->extFetch[b], wb.hits, r->dWork);
        else launch_closest<false>(ctx, s, sv.trace, q, 0, &wb.counters->extCount[b], &wb.counters->extFetch[b], wb.hits, nullptr);
        r->prof_end(s);
        r->launches++;
    };

    NX_CUDA(ctx, cudaMemsetAsync(wb.totals, 0, sizeof(WaveTotals), s));
    NX_CUDA(ctx, cudaEventRecord(r->evStart, s));
    r->launches = 0;
    for (uint32_t f = 0; f < nFrames; f++)
    {
        const uint32_t frame = firstFrame + f;
        NX_CUDA(ctx, cudaMemsetAsync(wb.counters, 0, sizeof(WaveCounters), s));
        r->prof_begin(0, s);
        nxi_launch_generate(sv, wb, frame, gGen, s);
        r->prof_end(s);
        r->launches++;
        closest(wb.ext[0], 0);
        if (r->queryPixel >= 0 && f == 0) {
            pixel_query_kernel<<<1, 1, 0, s>>>(wb.hits, pixel_to_slot((uint32_t)r->queryPixel, r->width, r->height), r->dQuery);
            r->queryPixel = -1;                                              // answered by this frame (pixelIdx = -1 in the reference)
        }
        for (uint32_t b = 1; b <= L; b++)
        {
            // two shadow queues: shade(b) refills queue b & 1, which the shadow trace of bounce b-2 must have consumed; the shadow
            // trace of bounce b-1 may still be running and overlaps this shade and the next extension trace
            if (b > 2) NX_CUDA(ctx, cudaStreamWaitEvent(s, r->evShadow[b & 1u], 0));
            r->prof_begin(2, s);
            nxi_launch_shade(sv, wb, b, frame, gShade, s);
            r->prof_end(s);
            r->launches++;
            NX_CUDA(ctx, cudaEventRecord(r->evShade, s));
            // The extension trace is on the critical path (the next shade waits for it), so it is launched first and its stream
            // has the higher priority; the shadow rays on the auxiliary stream fill in behind it and under the next shade
            // (the reference's graph runs the two traces as siblings).
            if (b < L) closest(wb.ext[b & 1u], b);
            NX_CUDA(ctx, cudaStreamWaitEvent(sa, r->evShade, 0));
            r->prof_begin(3, sa);
            if (work) launch_any<true>(ctx, sa, sv.trace, wb.shadow[b & 1u], 0, &wb.counters->shCount[b], &wb.counters->shFetch[b], nullptr, wb.shadowRad[b & 1u], wb.accum, r->dWork + 1);
            else launch_any<false>(ctx, sa, sv.trace, wb.shadow[b & 1u], 0, &wb.counters->shCount[b], &wb.counters->shFetch[b], nullptr, wb.shadowRad[b & 1u], wb.accum, nullptr);
            r->prof_end(sa);
            NX_CUDA(ctx, cudaEventRecord(r->evShadow[b & 1u], sa));
            r->launches++;
        }
        NX_CUDA(ctx, cudaStreamWaitEvent(s, r->evShadow[L & 1u], 0));      // the auxiliary stream is in order: the last shadow trace is the last to finish
        frame_totals_kernel<<<1, 32, 0, s>>>(wb, L);
        r->launches++;
    }
    NX_CUDA(ctx, cudaEventRecord(r->evStop, s));
    NX_CUDA(ctx, cudaGetLastError());
    r->frames += nFrames; r->timed = true; r->pathLengthLast = L;
    return NX_OK;
}

int nx_renderer_frame_count(const nx_renderer* r) { return r ? (int)r->frames : NX_ERR_INVALID; }

int nx_renderer_stats(nx_renderer* r, nx_frame_stats* out)
{
    if (!r || !out) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    std::memset(out, 0, sizeof(*out));
    NX_CUDA(ctx, cudaGetLastError());
    { const int rc = nxi_check_overflow(ctx); if (rc) return rc; }
    WaveTotals t{};
    NX_CUDA(ctx, cudaMemcpy(&t, r->wb.totals, sizeof(t), cudaMemcpyDeviceToHost));
    out->extension_rays = t.ext; out->shadow_rays = t.shadow; out->shaded_hits = t.shaded; out->frames = t.frames;
    if (r->timed) NX_CUDA(ctx, cudaEventElapsedTime(&out->device_ms, r->evStart, r->evStop));
    out->kernel_launches = r->launches;
    return NX_OK;
}

int nx_renderer_set_profiling(nx_renderer* r, int flags)
{
    if (!r) return NX_ERR_INVALID;
    r->profFlags = flags;
    return NX_OK;
}

int nx_renderer_profile(nx_renderer* r, nx_kernel_profile* out)
{
    if (!r || !out) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    std::memset(out, 0, sizeof(*out));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    for (const auto& l : r->evLaunches) {
        float ms = 0.f;
        NX_CUDA(ctx, cudaEventElapsedTime(&ms, r->evPool[l.second], r->evPool[l.second + 1]));
        out->ms[l.first] += ms; out->launches[l.first]++;
    }
    if (r->profFlags & 2) {
        TraceStats h[2];
        NX_CUDA(ctx, cudaMemcpy(h, r->dWork, sizeof(h), cudaMemcpyDeviceToHost));
        out->closest_work[0] = h[0].nodes; out->closest_work[1] = h[0].tris; out->closest_work[2] = h[0].insts; out->closest_work[3] = h[0].rays;
        out->any_work[0] = h[1].nodes; out->any_work[1] = h[1].tris; out->any_work[2] = h[1].insts; out->any_work[3] = h[1].rays;
        const unsigned long long* a = &h[0].iters; const unsigned long long* b = &h[1].iters;
        for (int k = 0; k < 10; k++) { out->closest_sched[k] = a[k]; out->any_sched[k] = b[k]; }
    }
    return NX_OK;
}

int nx_renderer_read_accum(nx_renderer* r, float* hostRgb)
{
    if (!r || !hostRgb) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    const size_t n = 3 * (size_t)r->width * r->height;
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    NX_CUDA(ctx, cudaMemcpyAsync(hostRgb, r->wb.accum, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const float inv = r->frames ? 1.0f / (float)r->frames : 0.f;
    for (size_t i = 0; i < n; i++) hostRgb[i] *= inv;
    return NX_OK;
}

int nx_renderer_accum_device(nx_renderer* r, float** outSum, uint32_t* outFrames)
{
    if (!r || !outSum) return NX_ERR_INVALID;
    *outSum = r->wb.accum; if (outFrames) *outFrames = r->frames;
    return NX_OK;
}
int nx_renderer_set_accum_frames(nx_renderer* r, uint32_t frames) { if (!r) return NX_ERR_INVALID; r->frames = frames; return NX_OK; }

int nx_display_transform(nx_ctx* ctx, const float* hostRgb, uint32_t count, int toneMapping, float exposure, uint32_t* hostRgba)
{
    if (!ctx || !hostRgb || !hostRgba || toneMapping < NX_TONE_NONE || toneMapping > NX_TONE_AGX_PUNCHY) return NX_ERR_INVALID;
    if (!count) return NX_OK;
    DeviceGuard guard(ctx->device);
    float* dIn = nullptr; uint32_t* dOut = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&dIn, 12 * (size_t)count, ctx->stream));
    NX_CUDA(ctx, cudaMallocAsync((void**)&dOut, 4 * (size_t)count, ctx->stream));
    NX_CUDA(ctx, cudaMemcpyAsync(dIn, hostRgb, 12 * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    nxi_launch_resolve(ctx->sm_count * 4, ctx->stream, dIn, count, 1.0f, exposure, toneMapping, dOut);
    NX_CUDA(ctx, cudaMemcpyAsync(hostRgba, dOut, 4 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(dIn, ctx->stream); cudaFreeAsync(dOut, ctx->stream);
    return NX_OK;
}

int nx_renderer_read_rgba8(nx_renderer* r, nx_scene* scene, uint32_t* hostRgba)
{
    if (!r || !scene || !hostRgba) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    const uint32_t count = r->width * r->height;
    uint32_t* d = nullptr;
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    NX_CUDA(ctx, cudaMallocAsync((void**)&d, 4 * (size_t)count, ctx->stream));
    nxi_launch_resolve(ctx->sm_count * 4, ctx->stream, r->wb.accum, count, r->frames ? 1.0f / (float)r->frames : 0.f, scene->settings.exposure, scene->settings.tone_mapping, d);
    NX_CUDA(ctx, cudaMemcpyAsync(hostRgba, d, 4 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(d, ctx->stream);
    return NX_OK;
}

// The reference's display path without the copy: PathTracer::Render hands AccumulateKernel the device pointer of a mapped OpenGL
// pixel buffer (src/OpenGL/PixelBuffer.cpp:4-40, src/Renderer/Renderer.cpp:41-48).  Here the caller supplies that pointer (W * H RGBA8
// words, row 0 = bottom row like a GL texture upload expects); the resolve is queued on the render stream behind the frame and the call
// returns at once - a viewer unmaps the buffer after synchronising nx_ctx_stream(), exactly where the reference calls cudaGraphicsUnmapResources.
int nx_renderer_present_device(nx_renderer* r, nx_scene* scene, uint32_t* devRgba)
{
    if (!r || !scene || !devRgba) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, devRgba) != cudaSuccess || (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        NX_FAIL(ctx, NX_ERR_INVALID, "PresentDevice: the destination is not device memory");
    }
    nxi_launch_resolve(ctx->sm_count * 4, ctx->stream, r->wb.accum, r->width * r->height, r->frames ? 1.0f / (float)r->frames : 0.f, scene->settings.exposure,
                       scene->settings.tone_mapping, devRgba);
    NX_CUDA(ctx, cudaGetLastError());
    return NX_OK;
}

int nx_renderer_present(nx_renderer* r, nx_scene* scene, uint32_t* hostRgba, int* outTicket)
{
    if (!r || !scene || !hostRgba || !outTicket) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    const uint32_t count = r->width * r->height;
    const int k = r->presentNext;
    cudaStream_t s = ctx->stream;
    if (!r->dRgba[k]) NX_CUDA(ctx, cudaMalloc((void**)&r->dRgba[k], 4 * (size_t)count));
    // nx_renderer_render leaves the render stream ordered after the shadow stream, so the accumulation is complete here.  The
    // device image of this slot may still be on its way to the host from two presents ago: the resolve waits for that copy.
    if (r->slotUsed[k]) NX_CUDA(ctx, cudaStreamWaitEvent(s, r->evCopied[k], 0));
    nxi_launch_resolve(ctx->sm_count * 4, s, r->wb.accum, count, r->frames ? 1.0f / (float)r->frames : 0.f, scene->settings.exposure, scene->settings.tone_mapping, r->dRgba[k]);
    NX_CUDA(ctx, cudaMemcpyAsync(r->hTotals + k, r->wb.totals, sizeof(WaveTotals), cudaMemcpyDeviceToHost, s));   // before the next render call clears them
    NX_CUDA(ctx, cudaEventRecord(r->evResolved[k], s));
    NX_CUDA(ctx, cudaStreamWaitEvent(r->copyStream, r->evResolved[k], 0));
    NX_CUDA(ctx, cudaMemcpyAsync(hostRgba, r->dRgba[k], 4 * (size_t)count, cudaMemcpyDeviceToHost, r->copyStream));
    NX_CUDA(ctx, cudaEventRecord(r->evCopied[k], r->copyStream));
    r->slotUsed[k] = true; r->presentNext = k ^ 1;
    *outTicket = k;
    return NX_OK;
}

int nx_renderer_present_wait(nx_renderer* r, int ticket, nx_frame_stats* out)
{
    if (!r || ticket < 0 || ticket > 1) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    if (!r->slotUsed[ticket]) NX_FAIL(ctx, NX_ERR_STATE, "nx_renderer_present_wait: ticket %d was never presented", ticket);
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaEventSynchronize(r->evCopied[ticket]));
    if (out) {
        std::memset(out, 0, sizeof(*out));
        const WaveTotals& t = r->hTotals[ticket];
        out->extension_rays = t.ext; out->shadow_rays = t.shadow; out->shaded_hits = t.shaded; out->frames = t.frames;
        out->kernel_launches = r->launches;
    }
    return NX_OK;
}

int nx_renderer_set_pixel_query(nx_renderer* r, uint32_t x, uint32_t y)
{
    if (!r) return NX_ERR_INVALID;
    if (x >= r->width || y >= r->height) NX_FAIL(r->ctx, NX_ERR_INVALID, "SetPixelQuery: pixel (%u, %u) outside %ux%u", x, y, r->width, r->height);
    r->queryPixel = (int64_t)y * r->width + x;      // PathTracer.cpp:236
    r->queryPending = true;
    return NX_OK;
}
int nx_renderer_pixel_query_pending(const nx_renderer* r) { return r ? (r->queryPending ? 1 : 0) : NX_ERR_INVALID; }
int nx_renderer_sync_pixel_query(nx_renderer* r, int32_t* outInstance)
{
    if (!r || !outInstance) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMemcpyAsync(r->hQuery, r->dQuery, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    r->queryPending = false;
    *outInstance = *r->hQuery;
    return NX_OK;
}

} // extern "C"
