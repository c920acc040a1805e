// wavequeue.cuh — the render loop as persistent warps fed by work queues in shared memory (included by integrator.cu).
//
// Same path state in shared memory, same scan / shade / regenerate code as the barrier-phased wavefront kernel
// (wavefront.cuh), but no CTA-wide phases: every slot sits in exactly one of five ring queues
//     SCAN | DIELECTRIC | NEWPATH (terminate + regenerate) | DIFFUSE | SPECULAR
// and each warp loops { take up to 32 (64 for SCAN) slots from the fullest queue; process them; push every slot to the
// queue of its next stage }.  ncu on the phased kernel shows 25-30 % of warp time parked at the end-of-iteration
// barrier, because a warp that drew dielectric chunks (exit search) takes 2.5x as long as one that drew diffuse chunks;
// with queues a slow work item delays nobody, the counting sort disappears, and class-specific fast paths (see the
// dielectric "creep" loop in path_shade) no longer unbalance anything.
//
// The queues are lock-free rings in shared memory: a producer reserves positions with one atomicAdd on the tail and
// then writes the slot ids; a consumer reserves a range with a CAS on the head and spins (briefly) on entries that are
// reserved but not yet written (the entry value itself is the ready flag; consumed entries are reset to EMPTY).  A
// slot id is in at most one queue, so WQ_SLOTS entries per ring never overflow.  (A first version used one CTA-wide
// spin lock: 9e9 lock spins per frame — the critical sections run at 1/32 of the SM's issue rate.)
#pragma once

namespace ptb {

#ifndef PTB_WQ_THREADS
#define PTB_WQ_THREADS 256
#endif
#ifndef PTB_WQ_SLOTS
#define PTB_WQ_SLOTS 512               // power of two; ~2x the slots that can be in flight (8 warps x 32..64)
#endif
#ifndef PTB_WQ_MIN_BLOCKS
#define PTB_WQ_MIN_BLOCKS 4
#endif
constexpr int WQ_THREADS = PTB_WQ_THREADS;
constexpr int WQ_WARPS = WQ_THREADS / 32;
constexpr int WQ_SLOTS = PTB_WQ_SLOTS;
constexpr int WQ_MASK = WQ_SLOTS - 1;
static_assert((WQ_SLOTS & WQ_MASK) == 0 && WQ_SLOTS % WQ_THREADS == 0, "PTB_WQ_SLOTS must be a power of two and a multiple of the CTA size");
enum : int { Q_SCAN = 0, Q_DIEL = 1, Q_NEWPATH = 2, Q_DIFFUSE = 3, Q_SPEC = 4, WQ_NQ = 5 };
__device__ __forceinline__ int queue_of_class(int cls) {       // CL_DIEL 0, CL_TERM 1, CL_REGEN 2, CL_DIFFUSE 3, CL_SPEC 4
    return (0x43221 >> (cls * 4)) & 15;
}

struct WqState : SlotState<WQ_SLOTS> {
    unsigned char cls[WQ_SLOTS];                 // class decided by the scan (or CL_REGEN after a terminated scatter)
    unsigned short ring[WQ_NQ][WQ_SLOTS];        // a slot is in at most one queue, so WQ_SLOTS entries never overflow
    int head[WQ_NQ], tail[WQ_NQ];                // monotonic; index = counter & WQ_MASK; atomics only
    int retired;                                 // slots that ran out of pixels; WQ_SLOTS = the CTA is done
};

#ifdef PTB_WQ_DEBUG
#define WQ_DBG(k, v) dbg[k] += (v)
#else
#define WQ_DBG(k, v)
#endif
constexpr unsigned short WQ_EMPTY = 0xFFFF;

template <bool STATS, bool MESH>
__global__ void __launch_bounds__(WQ_THREADS, PTB_WQ_MIN_BLOCKS)
integrate_wq_kernel(const __grid_constant__ FrameParams fp) {
    extern __shared__ uint4 s_raw[];
    WqState& S = *reinterpret_cast<WqState*>(s_raw);
    uint4* s_blob = s_raw + (sizeof(WqState) + 15) / 16;
    const int n_obj = c_scene.n_obj;
    {
        const int n_words = n_obj * 2 + c_scene.n_mat * 3;
        for (int i = threadIdx.x; i < n_words; i += blockDim.x) s_blob[i] = fp.scene_blob[i];
    }
    const DevObj* __restrict__ s_obj = reinterpret_cast<const DevObj*>(s_blob);
    const DevMat* __restrict__ s_mat = reinterpret_cast<const DevMat*>(s_blob + 2 * n_obj);

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_pix = fp.width * fp.height;
    const int n_box = c_scene.n_box;
    unsigned long long st[STATS ? kStatsWords : 1] = {0};
#ifdef PTB_WQ_DEBUG
    unsigned long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // scan items, shade items, scan slots, shade slots, idle polls, lock spins
#endif

    if (tid < WQ_NQ) { S.head[tid] = 0; S.tail[tid] = 0; }
    if (tid == 0) S.retired = 0;
    for (int i = tid; i < WQ_NQ * WQ_SLOTS; i += WQ_THREADS) (&S.ring[0][0])[i] = WQ_EMPTY;
    __syncthreads();
    // every slot gets its first camera ray and goes to the SCAN queue (slots that got no pixel are retired at once)
    for (int j = tid; j < WQ_SLOTS; j += WQ_THREADS) {
        S.pix[j] = -1; S.smp[j] = 0; S.depth[j] = 0;
        if (fp.max_depth > 0) path_regen<STATS>(S, fp, n_pix, j, false, st);
        if (S.pix[j] >= 0) S.ring[Q_SCAN][atomicAdd(&S.tail[Q_SCAN], 1) & WQ_MASK] = (unsigned short)j;
        else atomicAdd(&S.retired, 1);
    }
    __syncthreads();

    int patience = 8;
    for (;;) {
        // ------------------------------------------------------------ take work: the fullest queue
        int q = -1, n = 0, base = 0;
        if (lane == 0) {
            volatile int* head = S.head;
            volatile int* tail = S.tail;
            int bq = -1, bn = 0, bh = 0;
#pragma unroll
            for (int k = 0; k < WQ_NQ; ++k) {
                const int h = head[k], cnt = tail[k] - h;
                if (cnt > bn) { bn = cnt; bq = k; bh = h; }
            }
            // a full warp's worth — or, once this warp has polled in vain for a while, whatever there is
            if (bn >= 32 || (bn > 0 && patience <= 0)) {
                const int cap = bq == Q_SCAN ? 64 : 32;
                const int take = bn < cap ? bn : cap;
                if (atomicCAS(&S.head[bq], bh, bh + take) == bh) { q = bq; n = take; base = bh; }
            } else if (bn == 0 && *(volatile int*)&S.retired >= WQ_SLOTS) {
                q = -2;                                       // every slot is retired
            }
        }
        q = __reduce_max_sync(0xffffffffu, lane == 0 ? q : -3);
        n = __reduce_max_sync(0xffffffffu, lane == 0 ? n : 0);
        base = __reduce_max_sync(0xffffffffu, lane == 0 ? base : 0);
        if (q == -2) break;
        if (q < 0) { WQ_DBG(4, 1); --patience; __nanosleep(100); continue; }
        patience = 8;
        WQ_DBG(q == Q_SCAN ? 0 : 1, 1); WQ_DBG(q == Q_SCAN ? 2 : 3, n);
        if (STATS) { st[ST_LANE_TOTAL] += (q == Q_SCAN ? 2 : 1); }

        auto take_entry = [&](int pos) {                      // wait until the producer that reserved `pos` has written it
            volatile unsigned short* e = &S.ring[q][pos & WQ_MASK];
            unsigned short v;
            while ((v = *e) == WQ_EMPTY) { WQ_DBG(5, 1); }
            *e = WQ_EMPTY;
            return (int)v;
        };
        int slot0 = lane < n ? take_entry(base + lane) : -1;
        int slot1 = -1, dest0 = -1, dest1 = -1;

        if (q == Q_SCAN) {
            // -------------------------------------------------------- SCAN: up to two rays per lane against every object
            slot1 = lane + 32 < n ? take_entry(base + 32 + lane) : -1;
            __threadfence_block();
            RayK ray[2];
            float best[2] = {FLT_MAX, FLT_MAX};
            int bid[2] = {-1, -1};
            {
                const int j0 = slot0 >= 0 ? slot0 : 0, j1 = slot1 >= 0 ? slot1 : 0;     // idle lanes trace slot 0's ray
                ray[0] = make_ray(f3(S.ox[j0], S.oy[j0], S.oz[j0]), f3(S.dx[j0], S.dy[j0], S.dz[j0]));
                ray[1] = make_ray(f3(S.ox[j1], S.oy[j1], S.oz[j1]), f3(S.dx[j1], S.dy[j1], S.dz[j1]));
            }
            // (object records come from the shared-memory copy as 16-byte broadcast loads: the work choice above is data
            //  dependent, so ptxas does not keep this loop on the uniform datapath and constant-bank loads would be per-lane)
            const float4* __restrict__ s_objv = reinterpret_cast<const float4*>(s_blob);
#pragma unroll 2
            for (int i = 0; i < n_box; ++i) {
                const float4 lo = s_objv[2 * i], hi = s_objv[2 * i + 1];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float t;
                    if (hit_box(lo, hi, ray[k], 0.001f, best[k], t)) { best[k] = t; bid[k] = i; }
                }
            }
            for (int i = n_box; i < n_obj; ++i) {
                const float4 lo = s_objv[2 * i], hi = s_objv[2 * i + 1];
                const bool is_sphere = (__float_as_int(lo.w) & 3) == PTB_OBJ_SPHERE;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float t;
                    const bool h = is_sphere ? hit_sphere(lo, hi, ray[k], 0.001f, best[k], t) : hit_plane(lo, ray[k], 0.001f, best[k], t);
                    if (h) { best[k] = t; bid[k] = i; }
                }
            }
            if (MESH) {
                if (slot0 >= 0) bvh_closest<STATS>(fp.bvh_nodes, fp.bvh_tris, ray[0], 0.001f, best[0], bid[0], st);
                if (slot1 >= 0) bvh_closest<STATS>(fp.bvh_nodes, fp.bvh_tris, ray[1], 0.001f, best[1], bid[1], st);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = k == 0 ? slot0 : slot1;
                if (j < 0) continue;
                int c;
                if (bid[k] < 0) c = CL_TERM;
                else if (MESH && (bid[k] & kTriBit)) c = (__float_as_int(__ldg(fp.bvh_tris + 3 * (bid[k] & ~kTriBit) + 1).w) >> 3) & 7;
                else c = (s_obj[bid[k]].meta >> 3) & 7;
                S.best[j] = best[k]; S.bid[j] = bid[k]; S.cls[j] = (unsigned char)c;
                if (k == 0) dest0 = queue_of_class(c); else dest1 = queue_of_class(c);
                if (STATS) { st[ST_LANE_ACTIVE]++; st[ST_SEGMENTS]++; }
            }
        } else if (slot0 >= 0) {
            __threadfence_block();
            // -------------------------------------------------------- SHADE: 32 slots of one queue (= one class, or TERM + REGEN)
            const int j = slot0;
            path_shade<STATS, MESH>(S, fp, s_obj, s_mat, n_pix, j, (int)S.cls[j], st);
            if (STATS) st[ST_LANE_ACTIVE]++;
            if (S.pix[j] >= 0) {
                if (S.depth[j] > 0) dest0 = Q_SCAN;                       // scattered, or a fresh camera ray
                else { dest0 = Q_NEWPATH; S.cls[j] = (unsigned char)CL_REGEN; }   // path ended in the scatter: regenerate
            }
        }

        // ------------------------------------------------------------ push every slot to the queue of its next stage
        __threadfence_block();                                // slot state before the ring entry that publishes it
        const int n_retired = __popc(__ballot_sync(0xffffffffu, slot0 >= 0 && dest0 < 0));
#pragma unroll
        for (int d = 0; d < WQ_NQ; ++d) {
            const unsigned m0 = __ballot_sync(0xffffffffu, dest0 == d), m1 = __ballot_sync(0xffffffffu, dest1 == d);
            const int c0 = __popc(m0), c1 = __popc(m1);
            if (c0 + c1 == 0) continue;
            int t0 = 0;
            if (lane == 0) t0 = atomicAdd(&S.tail[d], c0 + c1);
            t0 = __shfl_sync(0xffffffffu, t0, 0);
            volatile unsigned short* ring = S.ring[d];
            if (dest0 == d) ring[(t0 + __popc(m0 & lt_mask)) & WQ_MASK] = (unsigned short)slot0;
            if (dest1 == d) ring[(t0 + c0 + __popc(m1 & lt_mask)) & WQ_MASK] = (unsigned short)slot1;
        }
        if (lane == 0 && n_retired) atomicAdd(&S.retired, n_retired);
    }

#ifdef PTB_WQ_DEBUG
    if (lane == 0 && fp.stats) for (int k = 0; k < 6; ++k) atomicAdd(fp.stats + kStatsWords + k, dbg[k]);
#endif
    if (STATS) {
        for (int k = 0; k < kStatsWords; ++k) {
            unsigned long long v = st[k];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0 && v) atomicAdd(fp.stats + k, v);
        }
    }
}

}  // namespace ptb
