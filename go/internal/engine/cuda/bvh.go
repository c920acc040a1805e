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
//	Nodes      16 float32 per node (64 bytes): both children's boxes as centre / half extent, then the two child links as bit
//	           patterns (link >= 0: inner node index; link < 0: leaf, ^link = firstTriangle<<2 | count-1)
//	Triangles  12 float32 per triangle (48 bytes), leaf order: v0 + bits(original index), e1 = v1-v0, e2 = v2-v0
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
	out.Nodes = append(out.Nodes, unsafe.Slice((*float32)(unsafe.Pointer(b.nodes)), int(b.info.n_nodes)*16)...)
	out.Triangles = append(out.Triangles, unsafe.Slice((*float32)(unsafe.Pointer(b.triangles)), int(b.info.n_triangles)*12)...)
	return out, nil
}
