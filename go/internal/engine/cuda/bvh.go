package cuda

/*
#include "ptb200.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// BVH is the flattened acceleration structure the CUDA backend traverses (north-star: "internal/scene gains a BVH builder that
// emits a flattened, cache-line-aligned node array").  The builder itself is the library's (binned SAH, ptb_bvh_build — the
// same code ptb_scene_upload runs), so the Go side and the device can never disagree about the layout:
//
//	Nodes      32 float32 per node (128 bytes, 4-wide): floats 6k..6k+5 = child k's box as centre / half extent (half < 0: unused),
//	           floats 24..27 = the child links as bit patterns (link >= 0: inner node index; link < 0: leaf,
//	           ^link = firstTriangle<<2 | count-1), floats 28..31 = 0
//	Triangles  16 float32 per triangle (64 bytes), leaf order: v0 + bits(original index), e1 = v1-v0, e2 = v2-v0, padding
type BVH struct {
	Nodes, Triangles []float32
	MaxDepth         int
	SAHCost, BuildMs float64
}

// BuildBVH builds the BVH over world-space triangles (9 float32 each, the output of scene.Object.WorldTriangles).  Rendering
// does not need it — Render hands the triangles to the library, which builds and caches the BVH per mesh generation — it is
// for tools that inspect, cache or serialise the structure.
func BuildBVH(tris []float32) (*BVH, error) {
	if len(tris) == 0 || len(tris)%9 != 0 {
		return nil, fmt.Errorf("BuildBVH: %d floats is not a whole number of triangles", len(tris))
	}
	var b C.ptb_bvh
	if rc := C.ptb_bvh_build((*C.float)(unsafe.Pointer(&tris[0])), C.int64_t(len(tris)/9), &b); rc != C.PTB_OK {
		return nil, fmt.Errorf("ptb_bvh_build: %s", C.GoString(C.ptb_last_error(nil)))
	}
	defer C.ptb_bvh_free(&b)
	out := &BVH{MaxDepth: int(b.info.max_depth), SAHCost: float64(b.info.sah_cost), BuildMs: float64(b.info.build_ms)}
	out.Nodes = append(out.Nodes, unsafe.Slice((*float32)(unsafe.Pointer(b.nodes)), int(b.info.n_nodes)*32)...)
	out.Triangles = append(out.Triangles, unsafe.Slice((*float32)(unsafe.Pointer(b.triangles)), int(b.info.n_triangles)*16)...)
	return out, nil
}
