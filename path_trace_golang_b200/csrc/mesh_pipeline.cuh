// mesh_pipeline.cuh — the render loop for scenes WITH triangle meshes as a wavefront across kernels (included by integrator.cu).
//
// Why not the persistent kernel of wavefront.cuh: BVH traversal is bound by the latency of dependent node fetches (L2 / HBM),
// so its throughput is (rays in flight) / (fetches per ray x latency).  Inside the persistent kernel the traversal is one
// phase of a CTA-wide loop — at any moment only about a third of an SM's warps are in it, and they wait for each other at
// the phase barriers (ncu r02d: issue slots 45 % busy, barrier 37 % + long-scoreboard 30 % of the warp stalls).  Here the
// traversal is a kernel of its own in which EVERY resident lane walks the BVH, refilling itself from a global ray queue, and
// the state of the paths lives in global memory between the kernels:
//
//   per iteration   mp_shade_scan_kernel   CTA b <-> path slots [512 b, 512 b + 512): load state + hit results -> classify ->
//                                          sort by class -> shade / regenerate (path_shade, path_regen of wavefront.cuh) ->
//                                          closest hit over the ANALYTIC objects for the new rays -> rays that reach the meshes'
//                                          bounds are appended to the ray queue -> store state
//                   mp_traverse_kernel     persistent warps: a lane without a ray takes the next queue entry, traverses the
//                                          4-wide BVH to the end (bvh.cuh), writes the hit back into the slot's result
//   the host launches iteration after iteration and reads an "any path alive" flag a few iterations behind (no bubble).
//
// Semantics are those of the persistent kernel: same helpers, same per-slot sample order, same counter-RNG draw order, so
// the same paths and the same per-pixel sums.  State traffic: 82 bytes per slot loaded + stored per iteration.
#pragma once

namespace ptb {

#ifndef PTB_MP_CTAS_PER_SM
#define PTB_MP_CTAS_PER_SM 8           // path slots in flight = SMs x this x WF_SLOTS
#endif
#ifndef PTB_MP_ROUND_NODES
#define PTB_MP_ROUND_NODES 4           // inner-node visits between two refills of a warp's idle lanes (traversal kernel)
#endif

// Control words (device): per iteration parity p: [4p] queue tail, [4p + 1] queue head; [8 + (i & 15)] "a path is alive after
// iteration i"; [24] iteration the next shade+scan kernel runs, [25] iteration of the traversal kernel in flight.  The kernels
// read the iteration number from these words, so the pair (shade+scan, traverse) is the same launch every time and the host
// replays it as a CUDA graph of several iterations (launching two kernels with 10 KB of parameters each costs more host time
// than they take to run).
constexpr int kMpCtlWords = 32;

__global__ void mp_init_kernel(MeshPool pool) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kMpCtlWords) pool.ctl[i] = 0u;
    if (i >= pool.n_slots) return;
    pool.A[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));      // no work item yet
    pool.B[i] = make_float4(1.f, 1.f, 1.f, __int_as_float(0));
    pool.O[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pool.D[i] = make_float4(0.f, 0.f, 1.f, 0.f);
    pool.best[i] = 0.0f; pool.bid[i] = -1;
    pool.dep[i] = 0;                                                 // "needs a camera ray": the first iteration regenerates every slot
}

template <bool STATS, bool BIG>
__global__ void __launch_bounds__(WF_THREADS, PTB_WF_MIN_BLOCKS)
mp_shade_scan_kernel(const __grid_constant__ KernelArgs ka, const __grid_constant__ MeshPool pool) {
    const FrameParams& fp = ka.fp;
    const SceneK& c_scene = ka.sc;
    extern __shared__ uint4 s_raw[];
    WfState& S = *reinterpret_cast<WfState*>(s_raw);
    uint4* s_blob = s_raw + (sizeof(WfState) + 15) / 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot0 = blockIdx.x * WF_SLOTS;
    const int iter = (int)pool.ctl[24];         // (written by the previous traversal kernel: stable while this kernel runs)
    const int par = iter & 1;
    if (blockIdx.x == 0 && tid == 0) {          // counters of the other parity: its traversal kernel has finished (stream order)
        pool.ctl[4 * (par ^ 1)] = 0u; pool.ctl[4 * (par ^ 1) + 1] = 0u;
        pool.ctl[8 + ((iter + 8) & 15)] = 0u;
        pool.ctl[25] = (unsigned)iter;
    }
    // ---- load.  A CTA whose slots are all retired has nothing to do (the frame is draining).
    unsigned short dp[WF_SPT];
    bool any = false;
#pragma unroll
    for (int k = 0; k < WF_SPT; ++k) { dp[k] = pool.dep[slot0 + tid + k * WF_THREADS]; any |= dp[k] != kDepDead; }
    if (!__syncthreads_or(any)) return;
    const int n_obj = c_scene.n_obj;
    if (!BIG) {
        const int n_words = n_obj * 2 + c_scene.n_mat * 3;
        for (int i = tid; i < n_words; i += blockDim.x) s_blob[i] = fp.scene_blob[i];
    }
    const DevObj* __restrict__ s_obj = BIG ? reinterpret_cast<const DevObj*>(fp.scene_blob) : reinterpret_cast<const DevObj*>(s_blob);
    const DevMat* __restrict__ s_mat = BIG ? reinterpret_cast<const DevMat*>(fp.scene_blob + 2 * n_obj) : reinterpret_cast<const DevMat*>(s_blob + 2 * n_obj);
    const int n_pix = fp.width * fp.rows;
    const ScanK sk(c_scene);
    unsigned long long st[STATS ? kStatsWords : 1] = {0};
    int cls[WF_SPT];
#pragma unroll
    for (int k = 0; k < WF_SPT; ++k) {
        const int j = tid + k * WF_THREADS, g = slot0 + j;
        S.O[j] = pool.O[g]; S.D[j] = pool.D[g]; S.B[j] = pool.B[g]; S.A[j] = pool.A[g];
        const float bt = pool.best[g];
        const int hb = pool.bid[g];
        S.best[j] = bt; S.bid[j] = hb; S.dep[j] = dp[k];
        int c;                                                         // class of the work the slot needs (wavefront.cuh)
        if (dp[k] == kDepDead) c = CL_DEAD;
        else if (dp[k] == 0u) c = PTB_MERGE_TERM_REGEN ? CL_TERM : CL_REGEN;
        else if (hb < 0) c = CL_TERM;
        else if (hb & kTriBit) c = (__float_as_int(__ldg(fp.bvh_tris + kTriQuads * (hb & ~kTriBit) + 1).w) >> 3) & 7;
        else c = (__ldg(&reinterpret_cast<const DevObj*>(fp.scene_blob)[hb].meta) >> 3) & 7;    // (the shared-memory copy is not complete yet)
        cls[k] = c;
        if (STATS) { st[ST_LANE_TOTAL]++; if (c != CL_DEAD && dp[k] != 0u) { st[ST_LANE_ACTIVE]++; st[ST_SEGMENTS]++; } }
    }
    if (tid == 0) S.n_list = 0;

    // ---- SORT: stable counting sort of the CTA's slots by class (see integrate_wf_kernel for the derivation)
    unsigned before[WF_SPT];
    unsigned mine = 0u;
    {
        unsigned Bm[WF_SPT][3];
#pragma unroll
        for (int k = 0; k < WF_SPT; ++k)
#pragma unroll
            for (int b = 0; b < 3; ++b) Bm[k][b] = __ballot_sync(0xffffffffu, (cls[k] >> b) & 1);
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < WF_SPT; ++k) {
            const unsigned n0 = (cls[k] & 1) ? 0u : ~0u, n1 = (cls[k] & 2) ? 0u : ~0u, n2 = (cls[k] & 4) ? 0u : ~0u;
            unsigned acc = 0u;
#pragma unroll
            for (int r = 0; r < k; ++r) acc += __popc((Bm[r][0] ^ n0) & (Bm[r][1] ^ n1) & (Bm[r][2] ^ n2));
            before[k] = acc + __popc((Bm[k][0] ^ n0) & (Bm[k][1] ^ n1) & (Bm[k][2] ^ n2) & lt);
        }
        const unsigned l0 = (lane & 1) ? 0u : ~0u, l1 = (lane & 2) ? 0u : ~0u, l2 = (lane & 4) ? 0u : ~0u;
#pragma unroll
        for (int r = 0; r < WF_SPT; ++r) mine += __popc((Bm[r][0] ^ l0) & (Bm[r][1] ^ l1) & (Bm[r][2] ^ l2));
    }
    if (lane < CL_COUNT) S.cnt[lane * WF_WARPS + warp] = (unsigned short)mine;
    __syncthreads();
    int n_dead;
    {
        constexpr int kPairs = CL_COUNT * WF_WARPS / 2;
        const unsigned pr = lane < kPairs ? reinterpret_cast<const unsigned*>(S.cnt)[lane] : 0u;
        const unsigned e0 = pr & 0xFFFFu, e1 = pr >> 16;
        unsigned inc = e0 + e1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned a = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += a;
        }
        const unsigned ex0 = inc - e0 - e1;
        const unsigned packed = ex0 | ((ex0 + e0) << 16);
        const int sh = (warp & 1) * 16;
#pragma unroll
        for (int k = 0; k < WF_SPT; ++k) {
            const int idx = cls[k] * WF_WARPS + warp;
            const int base = (int)((__shfl_sync(0xffffffffu, packed, idx >> 1) >> sh) & 0xFFFFu);
            S.perm[base + (int)before[k]] = (unsigned short)((tid + k * WF_THREADS) | (cls[k] << 12));
        }
        n_dead = WF_SLOTS - (int)(__shfl_sync(0xffffffffu, packed, CL_DEAD * WF_WARPS / 2) & 0xFFFFu);
    }
    if (tid == 0) S.shade_next = WF_WARPS;
    __syncthreads();
    const int live_chunks = (WF_SLOTS - n_dead + 31) >> 5;

    // ---- SHADE: one class per 32-slot chunk, heaviest class first, chunks handed out dynamically
#pragma unroll 1
    for (int chunk = warp; chunk < live_chunks;) {
        const unsigned pv = S.perm[chunk * 32 + lane];
        path_shade<STATS, true, BIG>(S, fp, c_scene, s_obj, s_mat, n_pix, pv & 0xFFF, pv >> 12, st);
        int nx = 0;
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(nx) : "r"((unsigned)__cvta_generic_to_shared(&S.shade_next)) : "memory");
        chunk = __shfl_sync(0xffffffffu, nx, 0);
    }
    __syncthreads();

    // ---- SCAN of the analytic world for the rays made above; rays that reach the meshes' bounds before their analytic hit are
    // queued for the traversal kernel (CTA-local list first: one global atomic per CTA)
    const float4 mc = make_float4(c_scene.mesh_c[0], c_scene.mesh_c[1], c_scene.mesh_c[2], 0.0f);
    const float4 mh = make_float4(c_scene.mesh_h[0], c_scene.mesh_h[1], c_scene.mesh_h[2], 0.0f);
    bool alive = false;
    static_assert(WF_SPT == WF_SG, "one scan group per thread");
    {
        RayK ray[WF_SG];
        float best[WF_SG];
        int bid[WF_SG];
        bool live[WF_SG];
#pragma unroll
        for (int k = 0; k < WF_SG; ++k) {
            const int j = tid + k * WF_THREADS;
            const float4 ov = S.O[j], dv = S.D[j];
            ray[k] = make_ray(f3(ov.x, ov.y, ov.z), f3(dv.x, dv.y, dv.z));
            best[k] = FLT_MAX; bid[k] = -1;
            const unsigned d2 = S.dep[j];
            live[k] = d2 != 0u && d2 != kDepDead;
            alive |= d2 != kDepDead;
        }
        scan_analytic<BIG, false>(c_scene, sk, s_obj, ray, best, bid);
#pragma unroll
        for (int k = 0; k < WF_SG; ++k) {
            const int j = tid + k * WF_THREADS;
            float tb;
            const bool need = live[k] && hit_box(mc, mh, ray[k], 0.001f, best[k], tb);
            S.best[j] = best[k]; S.bid[j] = bid[k];
            const unsigned m = __ballot_sync(0xffffffffu, need);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&S.n_list, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) S.perm[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;       // perm[] is free again
        }
    }
    alive = __syncthreads_or(alive);
    const int n_list = S.n_list;
    if (tid == 0) {
        S.next_chunk = n_list ? (int)atomicAdd(&pool.ctl[4 * par], (unsigned)n_list) : 0;      // this CTA's range of the global queue
        if (alive) pool.ctl[8 + (iter & 15)] = 1u;
    }
    __syncthreads();
    const int qbase = S.next_chunk;
    for (int i = tid; i < n_list; i += WF_THREADS) pool.queue[qbase + i] = slot0 + (int)S.perm[i];
    // ---- store
#pragma unroll
    for (int k = 0; k < WF_SPT; ++k) {
        const int j = tid + k * WF_THREADS, g = slot0 + j;
        pool.O[g] = S.O[j]; pool.D[g] = S.D[j]; pool.B[g] = S.B[j]; pool.A[g] = S.A[j];
        pool.best[g] = S.best[j]; pool.bid[g] = S.bid[j]; pool.dep[g] = S.dep[j];
    }
    if (STATS) {
        for (int k = 0; k < kStatsWords; ++k) {
            unsigned long long v = st[k];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0 && v) atomicAdd(fp.stats + k, v);
        }
    }
}

// Every resident lane walks the BVH; a lane whose ray is finished takes the next entry of the global queue.
#ifndef PTB_MP_TRAV_BLOCKS
#define PTB_MP_TRAV_BLOCKS 4           // resident 256-thread CTAs per SM the traversal kernel is compiled for (register budget)
#endif
template <bool STATS>
__global__ void __launch_bounds__(256, PTB_MP_TRAV_BLOCKS)
mp_traverse_kernel(const float4* __restrict__ nodes, const float4* __restrict__ tris, const MeshPool pool, unsigned long long* stats) {
    const int lane = threadIdx.x & 31;
    const int iter = (int)pool.ctl[25];
    const int par = iter & 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) pool.ctl[24] = (unsigned)(iter + 1);      // (nobody reads word 24 while this kernel runs)
    const int n_queue = (int)pool.ctl[4 * par];
    unsigned int* head = pool.ctl + 4 * par + 1;
    unsigned long long st[STATS ? kStatsWords : 1] = {0};
    int slot = -1;
    RayK tr;
    float tbest = 0.0f;
    int tbid = -1;
    TravState T;
    bool more = true;                  // the queue still has rays (warp-uniform)
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, slot < 0);
        if (idle && more) {
            int base = 0;
            if (lane == 0) base = (int)atomicAdd(head, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            more = base + __popc(idle) < n_queue;
            const int idx = base + __popc(idle & ((1u << lane) - 1u));
            if (slot < 0 && idx < n_queue) {
                slot = pool.queue[idx];
                const float4 ov = pool.O[slot], dv = pool.D[slot];
                tr = make_ray(f3(ov.x, ov.y, ov.z), f3(dv.x, dv.y, dv.z));
                tbest = pool.best[slot]; tbid = pool.bid[slot];
                trav_begin(T, nullptr, false, 0x7fffffff);
            }
        }
        if (__ballot_sync(0xffffffffu, slot >= 0) == 0u) break;
        if (slot >= 0) {
            if (trav_round<STATS>(nodes, tris, tr, 0.001f, tbest, tbid, st, T, PTB_MP_ROUND_NODES) == 1) {
                pool.best[slot] = tbest; pool.bid[slot] = tbid;
                slot = -1;
            }
        }
    }
    if (STATS) {
        for (int k = 0; k < kStatsWords; ++k) {
            unsigned long long v = st[k];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0 && v) atomicAdd(stats + k, v);
        }
    }
}

}  // namespace ptb
