// wavefront.cuh — the render loop as a persistent "wavefront in shared memory" kernel (included by integrator.cu).
//
// Why: in the pixel-per-lane megakernel (integrate_kernel) every lane of a warp executes the same closest-hit scan,
// but after it the lanes want different code (sky / emissive / lambert / mirror / glass + exit search / new
// camera ray), and ncu shows that part running with ~8 of 32 lanes active (profiles/r01b_*).  Here the path state
// lives in shared memory (WF_SLOTS = 512 slots per 256-thread CTA) and, after every scan, the CTA's slots are
// counting-sorted by the class of work they need, so that a warp shades 32 slots of the SAME class:
//
//   loop:  SCAN   thread i <-> slots i, i+256   closest hit over the world, two rays per record of the per-type scan
//                                               tables (warp-uniform loops, operands from the constant bank via LDCU)
//          SORT   ballots of the class bits -> in-warp ranks -> per-warp counts -> one shuffle prefix scan over
//                                               pair-packed counts -> stable permutation (2 barriers)
//          SHADE  warp w <-> 32-slot chunks w and 15-w of perm[]   scatter / terminate / regenerate, one class per
//                                               chunk; perm[] is sorted heaviest class first (serpentine pairing)
//
// A slot is bound to one pixel at a time and walks that pixel's samples in order with its fp32 sum in shared
// memory (deterministic per-pixel sum order, no atomics on radiance); when the pixel is done the slot writes it
// out and takes the next pixel index from a global atomic counter — dynamic scheduling at pixel granularity, so
// cheap sky pixels and expensive glass pixels balance across the whole chip.  On small frames a work item is a
// (pixel, sample sub-range) pair instead (FrameParams::split_k); with a row partition the items cover a subset of rows.
//
// Semantics are those of integrate_kernel (same helpers, same counter-RNG draw order), hence of the reference.
// MESH = true adds the BVH traversal of bvh.cuh to the scan (separate instantiation: the traversal code costs the
// mesh-free path 10 % if it is merely present): rays that reach the meshes' bounds are compacted CTA-wide, traversed by
// persistent lanes under a per-iteration step budget, and suspended (class CL_CONT) when the budget runs out.
#pragma once

namespace ptb {

#ifndef PTB_WF_THREADS
#define PTB_WF_THREADS 256
#endif
#ifndef PTB_WF_SPT
#define PTB_WF_SPT 2                   // path slots per thread: the scan tests two rays against each object record it loads
#endif
constexpr int WF_THREADS = PTB_WF_THREADS;
constexpr int WF_WARPS = WF_THREADS / 32;
constexpr int WF_SPT = PTB_WF_SPT;
#ifndef PTB_WF_SCAN_GROUP
#define PTB_WF_SCAN_GROUP 2           // rays tested together against each object record (register pressure)
#endif
constexpr int WF_SG = PTB_WF_SCAN_GROUP;
static_assert(WF_SPT % WF_SG == 0, "PTB_WF_SPT must be a multiple of PTB_WF_SCAN_GROUP");
constexpr int WF_SLOTS = WF_THREADS * WF_SPT;
constexpr int WF_CHUNKS = WF_SLOTS / 32;
// Classes in sort order = the order in which warps pull 32-slot chunks in the SHADE phase: heaviest first
// (longest-processing-time-first keeps the phase balanced), TERM and REGEN adjacent (they share the regeneration code).
// CL_DIEL..CL_SPEC match the class bits the host writes into DevObj::meta (api.cu).
// Measured chunk costs on C3 (cycles, -DPTB_WF_TIMING): DIEL 5500, TERM 3450, REGEN 2740, DIFFUSE 2210, SPEC 1630.
// CL_CONT (MESH only): the slot's BVH traversal ran out of its per-iteration step budget and continues next iteration.
#ifndef PTB_WF_DYNAMIC
#define PTB_WF_DYNAMIC 1              // SHADE phase: dynamic (1) or static serpentine (0) assignment of chunks to warps
#endif
#ifndef PTB_MERGE_TERM_REGEN
#define PTB_MERGE_TERM_REGEN 1        // sort slots that only need a new camera ray together with the terminating ones (one class boundary less)
#endif
enum : int { CL_DIEL = 0, CL_TERM = 1, CL_REGEN = 2, CL_DIFFUSE = 3, CL_SPEC = 4, CL_CONT = 5, CL_DEAD = 6, CL_COUNT = 7 };

// Path state, one entry per slot.  The SHADE phase reads and writes a slot through an arbitrary index (perm[]), so the state is
// packed into 16-byte vectors — one LDS.128 / STS.128 moves a whole vector and its rider — while the scalars the SCAN phase
// touches through its own index (stride 1: conflict-free) stay in arrays of their own.
constexpr unsigned short kDepDead = 0xFFFFu;   // dep[]: slot retired;  0 = no live path (needs a new camera ray);  else remaining depth
constexpr int kMaxDepth = 0xFFFE;
template <int N>
struct SlotState {
    float4 O[N];                       // ray origin            | RNG key (bits)
    float4 D[N];                       // ray direction         | RNG counter (bits)
    float4 B[N];                       // throughput beta       | sample index being traced (bits)
    float4 A[N];                       // pixel sum             | work item (pixel index, or plane * n_pix + pixel) (bits), -1 = none
    float best[N];                     // closest hit of the last scan ...
    int bid[N];                        // ... and its device object index (or kTriBit | triangle slot), -1 = miss
    unsigned short dep[N];             // remaining depth of the live path (renderer.go:287-289), 0 / kDepDead see above
};
struct WfState : SlotState<WF_SLOTS> {
    unsigned short perm[WF_SLOTS];     // slot | class << 12
    alignas(4) unsigned short cnt[CL_COUNT * WF_WARPS];
    unsigned char trav[WF_SLOTS];      // MESH: 1 = the slot's traversal is suspended (state in FrameParams::trav_scratch)
    int n_list, next_chunk;            // MESH: compacted list of slots to traverse (in perm[]) and its chunk dispenser
    int n_back;                        // MESH: rays starting their traversal are listed from the END of perm[] (resumed ones from the front)
    int shade_next;                    // PTB_WF_DYNAMIC: next unassigned 32-slot chunk of the SHADE phase
};

// Finish the slot's current sample and give it its next camera ray: next sample of the same pixel, or — when the
// pixel is complete — write the pixel out and take the next pixel from the global counter (renderer.go:171-221).
// Sample sub-range of global plane g of a split frame (FrameParams::split_*): boundaries of the WHOLE render.
__device__ __forceinline__ int plane_bound(const FrameParams& fp, int g) {
    return fp.full_begin + (int)((unsigned)(fp.full_end - fp.full_begin) * (unsigned)g / (unsigned)fp.split_total);
}

template <bool STATS, class SS>
__device__ __forceinline__ void path_regen(SS& S, const FrameParams& fp, const SceneK& sc, int n_pix, int j, bool sample_done, unsigned long long* st) {
    float4 av = S.A[j];
    int w = __float_as_int(av.w);                         // work item: the pixel, or plane * n_pix + pixel when the frame is split
    int s = __float_as_int(S.B[j].w) + (sample_done ? 1 : 0);
    const bool split = fp.planes != nullptr;             // (a launch may cover a single plane of a split render)
    int pix = w, plane = 0, s_end = fp.s_end;
    if (split && w >= 0) {
        plane = w / n_pix; pix = w - plane * n_pix;
        s_end = plane_bound(fp, fp.split_base + plane + 1);
    }
    if (w < 0 || s >= s_end) {
        if (w >= 0) {                                     // item complete: epilogue / accumulation buffer / partial-sum plane
            const float sx = av.x, sy = av.y, sz = av.z;
            if (split) {
                float* a = fp.planes + (size_t)w * 3; a[0] = sx; a[1] = sy; a[2] = sz;
            } else {
                if (fp.accum) { float* a = fp.accum + (size_t)pix * 3; a[0] = sx; a[1] = sy; a[2] = sz; }
                if (fp.rgba) {
                    const double inv_spp = 1.0 / (double)fp.spp_total;
                    reinterpret_cast<uchar4*>(fp.rgba)[pix] = make_uchar4(to_u8(sx, inv_spp), to_u8(sy, inv_spp), to_u8(sz, inv_spp), 255);
                }
            }
        }
        w = (int)atomicAdd(fp.work_counter, 1u);
        if (w >= n_pix * fp.split_k) { S.A[j] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1)); S.dep[j] = kDepDead; return; }
        pix = w; s = fp.s_begin;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        if (split) {
            plane = w / n_pix; pix = w - plane * n_pix;
            s = plane_bound(fp, fp.split_base + plane);
        } else if (fp.accum_resume) { const float* a = fp.accum + (size_t)pix * 3; a0 = a[0]; a1 = a[1]; a2 = a[2]; }
        S.A[j] = make_float4(a0, a1, a2, __int_as_float(w));
    }
    const int ly = pix / fp.width, px = pix - ly * fp.width;
    const int py = ly * fp.row_step + fp.row_offset;         // row partition: compact row ly is row py of the frame
    Rng rng;
    rng.key = fmix(fmix(fp.seed_key ^ (uint32_t)(py * fp.width + px)) + (uint32_t)s * kGolden);
    rng.ctr = 0u;
    const float u = ((float)px + rng.peek(0)) * fp.inv_w;                       // renderer.go:182
    const float v = ((fp.h_minus_1 - (float)py) + rng.peek(1)) * fp.inv_h;      // renderer.go:174,183
    rng.ctr = 2u;
    const DevCamera& cam = sc.cam;                                              // camera.go:60-74
    F3 dir = f3(cam.llc[0] + cam.horizontal[0] * u + cam.vertical[0] * v - cam.origin[0],
                cam.llc[1] + cam.horizontal[1] * u + cam.vertical[1] * v - cam.origin[1],
                cam.llc[2] + cam.horizontal[2] * u + cam.vertical[2] * v - cam.origin[2]);
    F3 org = f3(cam.origin[0], cam.origin[1], cam.origin[2]);
    if (cam.lens_radius > 0.0f) {
        F3 rd = in_unit_sphere(rng);
        float rx = rd.x * cam.lens_radius, ry = rd.y * cam.lens_radius;
        F3 off = f3(cam.u[0] * rx + cam.v[0] * ry, cam.u[1] * rx + cam.v[1] * ry, cam.u[2] * rx + cam.v[2] * ry);
        org = f3(org.x + off.x, org.y + off.y, org.z + off.z);
        dir = f3(dir.x - off.x, dir.y - off.y, dir.z - off.z);
    }
    sanitize_dir(dir);
    S.O[j] = make_float4(org.x, org.y, org.z, __uint_as_float(rng.key));
    S.D[j] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(rng.ctr));
    S.B[j] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(s));
    S.dep[j] = (unsigned short)fp.max_depth;
    if (STATS) st[ST_SAMPLES]++;
}

// One slot's work after a closest-hit scan, given its class c: scatter (+ exit search, Russian roulette), or add the
// sky / emitted radiance; terminate-class and regenerate-class slots then get their next camera ray.
// The scan table: kernel parameter (constant bank, uniform datapath) or, for BIG worlds, global memory.
template <bool BIG>
__device__ __forceinline__ float4 tab_ld4(const SceneK& sc, int i4) {
    if (BIG) return __ldg(sc.tab_global + i4);
    return reinterpret_cast<const float4*>(sc.scan_tab)[i4];
}
template <bool BIG>
__device__ __forceinline__ float tab_ld1(const SceneK& sc, int i) {
    if (BIG) return __ldg(reinterpret_cast<const float*>(sc.tab_global) + i);
    return sc.scan_tab[i];
}

template <bool STATS, bool MESH, bool BIG, class SS>
__device__ __forceinline__ void path_shade(SS& S, const FrameParams& fp, const SceneK& sc, const DevObj* __restrict__ s_obj, const DevMat* __restrict__ s_mat,
                                           int n_pix, int j, int c, unsigned long long* st) {
        if (c == CL_DIEL || c == CL_DIFFUSE || c == CL_SPEC) {
        const float4 ov = S.O[j], dv = S.D[j];
        const F3 ro = f3(ov.x, ov.y, ov.z);
        const F3 rd = f3(dv.x, dv.y, dv.z);
        const float t_hit = S.best[j];
        const int hid = S.bid[j];
        F3 p, n; bool front;
        int meta;
        if (MESH && (hid & kTriBit)) {
            tri_surface(fp.bvh_tris, hid & ~kTriBit, ro, rd, t_hit, p, n, front, meta);
            if (STATS) st[ST_ACC_MESH]++;
        } else {
            const DevObj ob = s_obj[hid];
            meta = ob.meta;
            if (STATS) st[ST_ACC_SPHERE + (meta & 3)]++;
            surface(ob, meta & 3, ro, rd, t_hit, p, n, front);
        }
        const DevMat m = s_mat[meta >> 6];
        Rng rng{__float_as_uint(ov.w), __float_as_uint(dv.w)};
        int depth = S.dep[j];

        uint32_t used = 0u;                               // draws consumed by this bounce (classes are warp-coherent: draw lazily)
        // unit incoming direction and its mirror image (math.go:39-46): every material but lambert needs them (a chunk of
        // lambert hits — most diffuse chunks — skips the block)
        float len = 1.0f, udn = 0.0f;
        F3 ud = rd, refl = rd;
        if (m.type != PTB_MAT_LAMBERT) {
            const float a = rd.x * rd.x + rd.y * rd.y + rd.z * rd.z;
            len = sqrt_(a);
            const float il = rcp_(len);
            ud = f3(rd.x * il, rd.y * il, rd.z * il);
            udn = ud.x * n.x + ud.y * n.y + ud.z * n.z;
            refl = f3(ud.x - n.x * 2.0f * udn, ud.y - n.y * 2.0f * udn, ud.z - n.z * 2.0f * udn);
        }

        F3 att = f3(m.albedo[0], m.albedo[1], m.albedo[2]), sd = refl, so = p;
        bool ok = true;
        if (m.type != PTB_MAT_LAMBERT && len == 0.0f) {   // materials.go:103-105, 178-180, 208-210
            ok = false;
            if (STATS) st[ST_END_NOSCATTER]++;
        } else if (c == CL_DIFFUSE) {                     // lambert (materials.go:76-97) / rough metal (:114-147)
            const F3 cd = cosine_direction(m.type == PTB_MAT_LAMBERT ? n : refl, rng.peek(0), rng.peek(1));
            used = 2u;
            if (m.type == PTB_MAT_LAMBERT) {
                sd = cd;
                if (m.rough > 1e-6f) {                    // rejection loop: consumes its draws itself (rare path)
                    rng.ctr += 2u;
                    F3 off = in_unit_sphere(rng);
                    used = 0u;
                    sd.x += off.x * m.rough * 0.1f; sd.y += off.y * m.rough * 0.1f; sd.z += off.z * m.rough * 0.1f;
                    sd = unit3(sd);
                }
            } else {
                const float alpha = m.rough * m.rough;
                float sx = refl.x * (1.0f - alpha) + cd.x * alpha;
                float sy = refl.y * (1.0f - alpha) + cd.y * alpha;
                float sz = refl.z * (1.0f - alpha) + cd.z * alpha;
                const float l2 = sx * sx + sy * sy + sz * sz;
                if (l2 < 1e-8f) { sx = refl.x; sy = refl.y; sz = refl.z; }
                else { const float i2 = rcp_(sqrt_(l2)); sx *= i2; sy *= i2; sz *= i2; }
                if (sx * n.x + sy * n.y + sz * n.z <= 0.0f) { sx = refl.x; sy = refl.y; sz = refl.z; }
                sd = f3(sx, sy, sz);
            }
        } else if (c == CL_DIEL) {                        // materials.go:162-200
            att = f3(1.0f, 1.0f, 1.0f);
            const float ratio = front ? rcp_(m.ior) : m.ior;
            const float cos_t = fminf(-udn, 1.0f);
            const float sin_t = sqrt_(1.0f - cos_t * cos_t);
            const bool cannot = ratio * sin_t > 1.0f;
            float r0 = (1.0f - ratio) * rcp_(1.0f + ratio);
            r0 = r0 * r0;
            const float om = 1.0f - cos_t;
            const float om2 = om * om;
            const float refl_prob = r0 + (1.0f - r0) * (om2 * om2 * om);   // Schlick, materials.go:226-231
            bool reflect = cannot;
            if (!cannot) { reflect = refl_prob > rng.peek(0); used = 1u; }  // `||` short-circuit: no draw when cannot
            if (!reflect) {                               // refractVec, math.go:48-64
                const float c2 = fminf(-ud.x * n.x - ud.y * n.y - ud.z * n.z, 1.0f);
                float qx = (ud.x + n.x * c2) * ratio, qy = (ud.y + n.y * c2) * ratio, qz = (ud.z + n.z * c2) * ratio;
                const float par = -sqrt_(fabsf(1.0f - (qx * qx + qy * qy + qz * qz)));
                sd = f3(qx + n.x * par, qy + n.y * par, qz + n.z * par);
            }
            sanitize_dir(sd);
            if (front) {                                  // exit search, renderer.go:316-371
                if (STATS) st[ST_EXIT_SCANS]++;
                const RayK er = make_ray(p, sd);
                float exit_t = FLT_MAX;
                bool hit_exit = false;
                F3 ep = p;
                // only dielectric objects can be accepted (:335): a back-face hit closer than the best so far whose squared
                // distance from p lies in (1e-8, 1000)
                auto consider = [&](float t, bool q_front, F3 q) {
                    if (!q_front && t < exit_t) {
                        float ex = q.x - p.x, ey = q.y - p.y, ez = q.z - p.z;
                        float d2 = ex * ex + ey * ey + ez * ez;
                        if (d2 > 1e-8f && d2 < 1000.0f) { hit_exit = true; exit_t = t; ep = q; }
                    }
                };
                if (sc.exit_typed) {                      // every dielectric object is a box or a sphere: typed records, no index loads
                    const int n_dbox = sc.n_dbox, n_dsph = sc.n_dsph;
                    const int dbox_off4 = sc.dbox_off4, dsph_off4 = sc.dsph_off4;
                    for (int k = 0; k < n_dbox; ++k) {
                        const float4 bc = tab_ld4<BIG>(sc, dbox_off4 + 2 * k), bh = tab_ld4<BIG>(sc, dbox_off4 + 2 * k + 1);
                        float t;
                        if (!hit_box(bc, bh, er, 0.0001f, exit_t, t)) continue;
                        DevObj eo;
                        eo.ax = bc.x; eo.ay = bc.y; eo.az = bc.z; eo.bx = bh.x; eo.by = bh.y; eo.bz = bh.z; eo.meta = PTB_OBJ_BOX; eo.world_idx = 0;
                        F3 q;
                        const bool qf = front_face_only(eo, PTB_OBJ_BOX, p, sd, t, q);
                        consider(t, qf, q);
                    }
                    for (int k = 0; k < n_dsph; ++k) {
                        const float4 sp = tab_ld4<BIG>(sc, dsph_off4 + k);
                        float t;
                        if (!hit_sphere4(sp.x, sp.y, sp.z, sp.w, er, 0.0001f, exit_t, t)) continue;
                        DevObj eo;
                        eo.ax = sp.x; eo.ay = sp.y; eo.az = sp.z; eo.bx = eo.by = eo.bz = 0.0f; eo.meta = PTB_OBJ_SPHERE; eo.world_idx = 0;
                        F3 q;
                        const bool qf = front_face_only(eo, PTB_OBJ_SPHERE, p, sd, t, q);
                        consider(t, qf, q);
                    }
                } else {
                    const int n_diel = sc.n_diel;
                    for (int k = 0; k < n_diel; ++k) {
                        const int ei = __ldg(sc.diel_idx + k);
                        const DevObj eo = s_obj[ei];
                        const int et = eo.meta & 3;
                        float t;
                        if (!hit_any(obj_lo(s_obj, ei), obj_hi(s_obj, ei), et, er, 0.0001f, exit_t, t)) continue;
                        F3 q;
                        const bool qf = front_face_only(eo, et, p, sd, t, q);
                        consider(t, qf, q);
                    }
                }
                if (hit_exit) {                           // renderer.go:352-369
                    float ex = ep.x - p.x, ey = ep.y - p.y, ez = ep.z - p.z;
                    float dist = sqrt_(ex * ex + ey * ey + ez * ez);
                    if (m.absorption[0] > 0.0f || m.absorption[1] > 0.0f || m.absorption[2] > 0.0f) {
                        att = f3(exp_(-m.absorption[0] * dist), exp_(-m.absorption[1] * dist), exp_(-m.absorption[2] * dist));
                    }
                    so = ep;
                }
            }
        }
        // (CL_SPEC — mirror and smooth metal: sd = refl, att = albedo, the defaults; materials.go:148-158, 205-221)

        bool done = !ok;
        if (ok) {
            if (STATS) st[ST_SCATTERS]++;
            if (depth <= 3) {                             // Russian roulette, renderer.go:374-393
                const float mx = fmaxf(att.x, fmaxf(att.y, att.z));
                if (mx < 1e-6f) {
                    done = true;
                } else {
                    const float pr = fminf(mx, 0.95f);
                    const float ur = rng.peek(used);
                    used += 1u;
                    if (ur > pr) done = true;
                    else { const float ip = rcp_(pr); att.x *= ip; att.y *= ip; att.z *= ip; }
                }
                if (STATS && done) st[ST_END_RR]++;
            }
            rng.ctr += used;
            if (!done) {                                  // renderer.go:398-403
                if (--depth <= 0) {                       // renderer.go:287-289
                    done = true;
                    if (STATS) st[ST_END_DEPTH]++;
                }
            }
        }
        if (done) {
            S.dep[j] = 0;                                 // regenerated next iteration, together with the other finished slots
        } else {
            float4 bv = S.B[j];
            bv.x *= att.x; bv.y *= att.y; bv.z *= att.z;
            S.B[j] = bv;
            if (c != CL_DIEL) sanitize_dir(sd);           // (dielectric: done before the exit search)
            S.O[j] = make_float4(so.x, so.y, so.z, ov.w);
            S.D[j] = make_float4(sd.x, sd.y, sd.z, __uint_as_float(rng.ctr));
            S.dep[j] = (unsigned short)depth;
        }
    } else if (c == CL_TERM && (!PTB_MERGE_TERM_REGEN || S.dep[j] != 0)) {     // sky (renderer.go:304-306) or emissive hit (:308-312)
        F3 e;                                             // (merged classes: a slot without a live path only regenerates)
        const int hb = S.bid[j];
        if (hb < 0) {
            const float4 dv = S.D[j];
            e = sky_color(sc.sky, f3(dv.x, dv.y, dv.z));
            if (STATS) st[ST_END_SKY]++;
        } else {
            const bool is_tri = MESH && (hb & kTriBit) != 0;
            const int meta = is_tri ? __float_as_int(__ldg(fp.bvh_tris + kTriQuads * (hb & ~kTriBit) + 1).w) : s_obj[hb].meta;
            const DevMat& m = s_mat[meta >> 6];
            e = f3(m.emit[0], m.emit[1], m.emit[2]);
            if (STATS) { st[ST_END_EMISSIVE]++; st[is_tri ? ST_ACC_MESH : ST_ACC_SPHERE + (meta & 3)]++; }
        }
        const float4 bv = S.B[j];
        float4 av = S.A[j];
        av.x += bv.x * e.x; av.y += bv.y * e.y; av.z += bv.z * e.z;
        S.A[j] = av;
    }
    if (c == CL_TERM || c == CL_REGEN) path_regen<STATS>(S, fp, sc, n_pix, j, true, st);
}

// Loop bounds and table offsets of the closest-hit scan, read once per kernel from the scene header (warp-uniform values).
struct ScanK {
    int n_obj, n_box, n_box_groups, n_plane_run, n_sphere_groups, plane_off4, sphere_off4, n_typed, sphere_base;
    __device__ __forceinline__ explicit ScanK(const SceneK& sc)
        : n_obj(sc.n_obj), n_box(sc.n_box), n_box_groups(sc.n_box_groups), n_plane_run(sc.n_plane_run), n_sphere_groups(sc.n_sphere_groups),
          plane_off4(sc.plane_off4), sphere_off4(sc.sphere_off4), n_typed(sc.n_typed), sphere_base(sc.n_box + sc.n_plane_run) {}
};

// Closest hit of WF_SG rays over the analytic world (renderer.go:292-302): every record of the per-type scan tables is tested
// against all the rays as it is loaded (warp-uniform loops; operands from the constant bank via LDCU unless BIG).
template <bool BIG, bool PK>
__device__ __forceinline__ void scan_analytic(const SceneK& sc, const ScanK& k_, const DevObj* __restrict__ s_obj, const RayK (&ray)[WF_SG],
                                              float (&best)[WF_SG], int (&bid)[WF_SG]) {
    for (int gi = 0; gi < k_.n_box_groups; ++gi) {              // kBoxGroup boxes per trip, 6 floats each (scene_dev.h)
        const int q = gi * (kBoxGroup * 6 / 4);
        float bx[kBoxGroup * 6];
#pragma unroll
        for (int v = 0; v < kBoxGroup * 6 / 4; ++v) { const float4 w = tab_ld4<BIG>(sc, q + v); bx[4 * v] = w.x; bx[4 * v + 1] = w.y; bx[4 * v + 2] = w.z; bx[4 * v + 3] = w.w; }
#pragma unroll
        for (int u = 0; u < kBoxGroup; ++u) {
            const float4 lo = make_float4(bx[6 * u], bx[6 * u + 1], bx[6 * u + 2], 0.0f);
            const float4 hi = make_float4(bx[6 * u + 3], bx[6 * u + 4], bx[6 * u + 5], 0.0f);
#pragma unroll
            for (int k = 0; k < WF_SG; ++k) {
                float t;
                if (hit_box(lo, hi, ray[k], 0.001f, best[k], t)) { best[k] = t; bid[k] = gi * kBoxGroup + u; }
            }
        }
    }
    for (int i = 0; i < k_.n_plane_run; ++i) {
        const float py = tab_ld1<BIG>(sc, k_.plane_off4 * 4 + i);
#pragma unroll
        for (int k = 0; k < WF_SG; ++k) {
            float t;
            if (hit_plane1(py, ray[k], 0.001f, best[k], t)) { best[k] = t; bid[k] = k_.n_box + i; }
        }
    }
    if constexpr (PK) {                                         // the two rays of the thread per instruction (integrator.cu "two rays per instruction")
        static_assert(WF_SG == 2, "the packed sphere test pairs the two rays of a thread");
        const RayP rp = pair_rays(ray[0], ray[1]);
        for (int gi = 0; gi < k_.n_sphere_groups; ++gi) {
#pragma unroll
            for (int u = 0; u < kSphereGroup; ++u) {
                const float4 sp = tab_ld4<BIG>(sc, k_.sphere_off4 + gi * kSphereGroup + u);
                float t[2]; bool h[2];
                hit_sphere_pair(sp.x, sp.y, sp.z, sp.w, rp, 0.001f, best, t, h);
#pragma unroll
                for (int k = 0; k < 2; ++k) if (h[k]) { best[k] = t[k]; bid[k] = k_.sphere_base + gi * kSphereGroup + u; }
            }
        }
    } else {
        for (int gi = 0; gi < k_.n_sphere_groups; ++gi) {
#pragma unroll
            for (int u = 0; u < kSphereGroup; ++u) {
                const float4 sp = tab_ld4<BIG>(sc, k_.sphere_off4 + gi * kSphereGroup + u);
#pragma unroll
                for (int k = 0; k < WF_SG; ++k) {
                    float t;
                    if (hit_sphere4(sp.x, sp.y, sp.z, sp.w, ray[k], 0.001f, best[k], t)) { best[k] = t; bid[k] = k_.sphere_base + gi * kSphereGroup + u; }
                }
            }
        }
    }
    for (int i = k_.n_typed; i < k_.n_obj; ++i) {                  // whatever follows the typed runs in world order
        const float4 lo = obj_lo(s_obj, i), hi = obj_hi(s_obj, i);
        const bool is_sphere = (__float_as_int(lo.w) & 3) == PTB_OBJ_SPHERE;
#pragma unroll
        for (int k = 0; k < WF_SG; ++k) {
            float t;
            const bool h = is_sphere ? hit_sphere(lo, hi, ray[k], 0.001f, best[k], t) : hit_plane(lo, ray[k], 0.001f, best[k], t);
            if (h) { best[k] = t; bid[k] = i; }
        }
    }
}

template <bool STATS, bool MESH, bool BIG, bool PK>
__global__ void __launch_bounds__(WF_THREADS, PTB_WF_MIN_BLOCKS)
integrate_wf_kernel(const __grid_constant__ KernelArgs ka) {
    const FrameParams& fp = ka.fp;
    const SceneK& c_scene = ka.sc;
    extern __shared__ uint4 s_raw[];
    WfState& S = *reinterpret_cast<WfState*>(s_raw);
    uint4* s_blob = s_raw + (sizeof(WfState) + 15) / 16;
    const int n_obj = c_scene.n_obj;
    if (!BIG) {
        const int n_words = n_obj * 2 + c_scene.n_mat * 3;
        for (int i = threadIdx.x; i < n_words; i += blockDim.x) s_blob[i] = fp.scene_blob[i];
    }
    // object / material records for the divergent look-ups: the CTA's shared-memory copy, or (BIG) global memory in place
    const DevObj* __restrict__ s_obj = BIG ? reinterpret_cast<const DevObj*>(fp.scene_blob) : reinterpret_cast<const DevObj*>(s_blob);
    const DevMat* __restrict__ s_mat = BIG ? reinterpret_cast<const DevMat*>(fp.scene_blob + 2 * n_obj) : reinterpret_cast<const DevMat*>(s_blob + 2 * n_obj);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_pix = fp.width * fp.rows;
    const ScanK sk(c_scene);
    unsigned long long st[STATS ? kStatsWords : 1] = {0};

#pragma unroll
    for (int k = 0; k < WF_SPT; ++k) {
        const int j = tid + k * WF_THREADS;
        S.A[j] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); S.B[j] = make_float4(1.f, 1.f, 1.f, __int_as_float(0));
        S.O[j] = make_float4(0.f, 0.f, 0.f, 0.f); S.D[j] = make_float4(0.f, 0.f, 1.f, 0.f);
        S.dep[j] = kDepDead; S.trav[j] = 0;
        if (fp.max_depth > 0) path_regen<STATS>(S, fp, c_scene, n_pix, j, false, st);
    }
    if (tid == 0) { S.n_list = 0; S.next_chunk = 0; S.n_back = 0; }
    __syncthreads();

#ifdef PTB_WF_TIMING
    // tm: 0 scan, 1 sort (both halves), 2 shade (own), 3 shade (CTA max, thread 0 only), 4 iterations
    long long tm[6] = {0, 0, 0, 0, 0, 0};
    long long tq = clock64();
    __shared__ int s_tmax;
    if (tid == 0) s_tmax = 0;
#define PTB_TICK(k) { long long now_ = clock64(); tm[k] += now_ - tq; tq = now_; }
#define PTB_MARK() { tq = clock64(); }
#else
#define PTB_TICK(k)
#define PTB_MARK()
#endif
    for (;;) {
        // ------------------------------------------------------------ SCAN (thread <-> its own WF_SPT slots, WF_SG rays at a time)
        int cls[WF_SPT];
#pragma unroll
        for (int g = 0; g < WF_SPT; g += WF_SG) {
            RayK ray[WF_SG];
            float best[WF_SG];
            int bid[WF_SG];
#pragma unroll
            for (int k = 0; k < WF_SG; ++k) {
                const int j = tid + (g + k) * WF_THREADS;
                const float4 ov = S.O[j], dv = S.D[j];
                ray[k] = make_ray(f3(ov.x, ov.y, ov.z), f3(dv.x, dv.y, dv.z));
                best[k] = FLT_MAX; bid[k] = -1;
            }
            scan_analytic<BIG, PK>(c_scene, sk, s_obj, ray, best, bid);
            if (MESH) {
                // EXTENSION: triangle meshes, tested after the analytic objects (bvh.cuh).  Only rays that reach the meshes'
                // bounds before their analytic hit (or whose traversal was suspended) are traversed, and they are compacted
                // CTA-wide first: a warp walks the BVH with 32 such rays instead of the few its own slots happen to hold.
                static_assert(WF_SPT == WF_SG, "the mesh path compacts all of a thread's slots at once");
                const float4 mc = make_float4(c_scene.mesh_c[0], c_scene.mesh_c[1], c_scene.mesh_c[2], 0.0f);
                const float4 mh = make_float4(c_scene.mesh_h[0], c_scene.mesh_h[1], c_scene.mesh_h[2], 0.0f);
#pragma unroll
                for (int k = 0; k < WF_SG; ++k) {
                    const int j = tid + (g + k) * WF_THREADS;
                    const bool live = S.dep[j] != 0 && S.dep[j] != kDepDead;
                    const bool resume = live && S.trav[j] != 0;
                    float tb;
                    const bool need = resume || (live && hit_box(mc, mh, ray[k], 0.001f, best[k], tb));
                    if (!resume) { S.best[j] = best[k]; S.bid[j] = bid[k]; }     // a suspended slot keeps its partial result
                    // Longest-processing-time-first: a ray whose traversal was suspended has already used up one step budget, so it
                    // is likely to be long — it goes to the FRONT of the list (perm[0..)), rays that start now to the back
                    // (perm[WF_SLOTS-1] downwards), and the lanes take the list front to back: the long rays start first and the
                    // phase does not end on a late, long straggler.
                    const unsigned mr = __ballot_sync(0xffffffffu, resume), mn = __ballot_sync(0xffffffffu, need && !resume);
                    int base_r = 0, base_n = 0;
                    if (lane == 0 && mr) base_r = atomicAdd(&S.n_list, __popc(mr));
                    if (lane == 0 && mn) base_n = atomicAdd(&S.n_back, __popc(mn));
                    base_r = __shfl_sync(0xffffffffu, base_r, 0); base_n = __shfl_sync(0xffffffffu, base_n, 0);
                    const unsigned lt = (1u << lane) - 1u;
                    if (resume) S.perm[base_r + __popc(mr & lt)] = (unsigned short)j;                // perm[] is free until the sort
                    else if (need) S.perm[WF_SLOTS - 1 - (base_n + __popc(mn & lt))] = (unsigned short)j;
                }
                __syncthreads();
                const int n_front = S.n_list, n_list = n_front + S.n_back;
                {   // persistent lanes: a lane whose ray finished (or ran out of budget) takes the next ray of the list
                    int tj = -1;                     // slot this lane is traversing, -1 = none
                    RayK tr;
                    float tbest = 0.0f;
                    int tbid = -1;
                    TravState T;
                    bool more = true;                // the list still has rays (warp-uniform)
                    for (;;) {
                        const unsigned idle = __ballot_sync(0xffffffffu, tj < 0);
                        if (idle && more) {
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&S.next_chunk, __popc(idle));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            more = base + __popc(idle) < n_list;
                            const int idx = base + __popc(idle & ((1u << lane) - 1u));
                            if (tj < 0 && idx < n_list) {
                                tj = S.perm[idx < n_front ? idx : WF_SLOTS - 1 - (idx - n_front)];
                                const float4 ov = S.O[tj], dv = S.D[tj];
                                tr = make_ray(f3(ov.x, ov.y, ov.z), f3(dv.x, dv.y, dv.z));
                                tbest = S.best[tj]; tbid = S.bid[tj];
                                trav_begin(T, fp.trav_scratch + ((size_t)blockIdx.x * WF_SLOTS + tj) * kTravStride, S.trav[tj] != 0, PTB_BVH_STEP_BUDGET);
                            }
                        }
                        if (__ballot_sync(0xffffffffu, tj >= 0) == 0u) break;
#ifdef PTB_BVH_TAIL_BUDGET
                        if (!more && T.budget > PTB_BVH_TAIL_BUDGET) T.budget = PTB_BVH_TAIL_BUDGET;   // list exhausted: do not let the CTA wait for stragglers
#endif
                        if (tj >= 0) {
                            const int status = trav_round<STATS>(fp.bvh_nodes, fp.bvh_tris, tr, 0.001f, tbest, tbid, st, T, PTB_BVH_ROUND_NODES);
                            if (status != 0) {
                                if (status == 2) trav_save(T, fp.trav_scratch + ((size_t)blockIdx.x * WF_SLOTS + tj) * kTravStride);
                                S.best[tj] = tbest; S.bid[tj] = tbid; S.trav[tj] = status == 2 ? 1 : 0;
                                tj = -1;
                            }
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) { S.n_list = 0; S.next_chunk = 0; S.n_back = 0; }  // next use is behind the sort's barriers
#pragma unroll
                for (int k = 0; k < WF_SG; ++k) {
                    const int j = tid + (g + k) * WF_THREADS;
                    best[k] = S.best[j]; bid[k] = S.bid[j];
                }
            }
#pragma unroll
            for (int k = 0; k < WF_SG; ++k) {
                const int j = tid + (g + k) * WF_THREADS;
                int c;
                const unsigned dp = S.dep[j];
                if (dp == kDepDead) c = CL_DEAD;
                else if (dp == 0u) c = PTB_MERGE_TERM_REGEN ? CL_TERM : CL_REGEN;
                else if (MESH && S.trav[j] != 0) c = CL_CONT;
                else if (bid[k] < 0) c = CL_TERM;
                else if (MESH && (bid[k] & kTriBit)) c = (__float_as_int(__ldg(fp.bvh_tris + kTriQuads * (bid[k] & ~kTriBit) + 1).w) >> 3) & 7;
                else c = (s_obj[bid[k]].meta >> 3) & 7;
                cls[g + k] = c;
                S.best[j] = best[k]; S.bid[j] = bid[k];
                if (STATS) { st[ST_LANE_TOTAL]++; if (c != CL_DEAD && dp != 0u) { st[ST_LANE_ACTIVE]++; if (c != CL_CONT) st[ST_SEGMENTS]++; } }
            }
        }

        PTB_TICK(0)
        // ------------------------------------------------------------ SORT (stable counting sort of the CTA's slots by class)
        // In-warp ranks from ballots of the three class BITS (not one ballot per class): the lanes whose row-r class
        // equals class c are AND_b (B[r][b] ^ nmask_b(c)), nmask_b(c) = 0 if bit b of c is set, else ~0.
        unsigned before[WF_SPT];            // slots of the same class that precede mine inside this warp
        unsigned mine = 0u;                 // lane c < CL_COUNT: this warp's number of class-c slots
        {
            unsigned B[WF_SPT][3];
#pragma unroll
            for (int k = 0; k < WF_SPT; ++k)
#pragma unroll
                for (int b = 0; b < 3; ++b) B[k][b] = __ballot_sync(0xffffffffu, (cls[k] >> b) & 1);
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < WF_SPT; ++k) {
                const unsigned n0 = (cls[k] & 1) ? 0u : ~0u, n1 = (cls[k] & 2) ? 0u : ~0u, n2 = (cls[k] & 4) ? 0u : ~0u;
                unsigned acc = 0u;
#pragma unroll
                for (int r = 0; r < k; ++r) acc += __popc((B[r][0] ^ n0) & (B[r][1] ^ n1) & (B[r][2] ^ n2));
                before[k] = acc + __popc((B[k][0] ^ n0) & (B[k][1] ^ n1) & (B[k][2] ^ n2) & lt);
            }
            const unsigned l0 = (lane & 1) ? 0u : ~0u, l1 = (lane & 2) ? 0u : ~0u, l2 = (lane & 4) ? 0u : ~0u;
#pragma unroll
            for (int r = 0; r < WF_SPT; ++r) mine += __popc((B[r][0] ^ l0) & (B[r][1] ^ l1) & (B[r][2] ^ l2));
        }
        if (lane < CL_COUNT) S.cnt[lane * WF_WARPS + warp] = (unsigned short)mine;
        PTB_TICK(1)
        __syncthreads();
        PTB_MARK()
        // exclusive prefix over the CL_COUNT x WF_WARPS (class-major, warp-minor) counts, redundantly in every warp: lane l
        // owns entries 2l and 2l+1 (one 32-bit load of two 16-bit counts), one 5-step scan of the pair sums
        static_assert(CL_COUNT * WF_WARPS <= 64 && WF_WARPS % 2 == 0 && WF_SLOTS < 65536, "pair-packed prefix");
        int base[WF_SPT];
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
            const unsigned packed = ex0 | ((ex0 + e0) << 16);      // exclusive prefixes of entries 2l (low) and 2l+1 (high)
            const int sh = (warp & 1) * 16;                           // entry index = class * WF_WARPS + warp: its parity is the warp's
#pragma unroll
            for (int k = 0; k < WF_SPT; ++k) {
                const int idx = cls[k] * WF_WARPS + warp;
                base[k] = (int)((__shfl_sync(0xffffffffu, packed, idx >> 1) >> sh) & 0xFFFFu);
            }
            // live slots = everything sorted before the CL_DEAD row
            n_dead = WF_SLOTS - (int)(__shfl_sync(0xffffffffu, packed, CL_DEAD * WF_WARPS / 2) & 0xFFFFu);
        }
#pragma unroll
        for (int k = 0; k < WF_SPT; ++k)
            S.perm[base[k] + (int)before[k]] = (unsigned short)((tid + k * WF_THREADS) | (cls[k] << 12));
#if PTB_WF_DYNAMIC
        if (tid == 0) S.shade_next = WF_WARPS;      // (the previous SHADE phase ended behind a barrier; the next starts behind the one below)
#endif
        PTB_TICK(1)
        __syncthreads();
        PTB_MARK()
        if (n_dead == WF_SLOTS) break;                                                // every slot retired (CTA-uniform)
        const int live_chunks = (WF_SLOTS - n_dead + 31) >> 5;                         // CL_DEAD sorts last

        // ------------------------------------------------------------ SHADE (one class per 32-slot chunk of perm[])
        // Static, serpentine chunk assignment: perm[] is sorted heaviest class first, so warp w takes chunks w,
        // 2W-1-w, 2W+w, ... and pairs a heavy chunk with a light one (keeps the phase balanced without atomics,
        // and keeps the loop structure provably uniform so the scan above stays on the uniform datapath).
#if PTB_WF_DYNAMIC
        // Dynamic chunk assignment: perm[] is sorted heaviest class first; every warp starts with chunk `warp` and then takes
        // the next unassigned chunk (one shared-memory atomic by lane 0, broadcast by shuffle — a warp-uniform value, so the
        // loop structure stays provably uniform): longest-processing-time-first list scheduling instead of a fixed pairing.
#pragma unroll 1
        for (int chunk = warp; chunk < live_chunks;) {
#else
#pragma unroll 1
        for (int q = 0; q < WF_SPT; ++q) {
        const int chunk = (q & 1) ? (q + 1) * WF_WARPS - 1 - warp : q * WF_WARPS + warp;
        if (chunk >= live_chunks) continue;
#endif
        const unsigned pv = S.perm[chunk * 32 + lane];
        const int j = pv & 0xFFF, c = pv >> 12;
#ifdef PTB_WF_TIMING
        const long long tc0 = clock64();
        const int c_lane0 = __shfl_sync(0xffffffffu, c, 0), c_lane31 = __shfl_sync(0xffffffffu, c, 31);
#endif
        path_shade<STATS, MESH, BIG>(S, fp, c_scene, s_obj, s_mat, n_pix, j, c, st);
#ifdef PTB_WF_TIMING
        if (lane == 0 && fp.stats && c_lane0 == c_lane31) {      // homogeneous chunks only
            atomicAdd(fp.stats + kStatsWords + 8 + 2 * c_lane0, (unsigned long long)(clock64() - tc0));
            atomicAdd(fp.stats + kStatsWords + 8 + 2 * c_lane0 + 1, 1ull);
        }
#endif
#if PTB_WF_DYNAMIC
        {   // (inline PTX: the atomicAdd intrinsic wraps a one-lane atomic in ~10 instructions of warp aggregation)
            int nx = 0;
            if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(nx) : "r"((unsigned)__cvta_generic_to_shared(&S.shade_next)) : "memory");
            chunk = __shfl_sync(0xffffffffu, nx, 0);
        }
#endif
        }   // chunk loop
#ifdef PTB_WF_TIMING
        { long long now_ = clock64(); int dt_ = (int)(now_ - tq); tm[2] += dt_; if (lane == 0) atomicMax(&s_tmax, dt_); }
#endif
        __syncthreads();
#ifdef PTB_WF_TIMING
        if (tid == 0) { tm[3] += s_tmax; tm[4] += 1; }
        __syncthreads();
        if (tid == 0) s_tmax = 0;
#endif
        PTB_MARK()
    }

#ifdef PTB_WF_TIMING
    if (lane == 0 && fp.stats) for (int k = 0; k < 6; ++k) atomicAdd(fp.stats + kStatsWords + k, (unsigned long long)tm[k]);
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
