// Package cuda is the reference-side binding of libptb200.so: the B200 CUDA backend for the per-pixel render
// loop.  It sits exactly where internal/engine/gpu sits today (gpu.Render, gpu.go:2534) and has the same
// signature, so engine.RenderInto only needs one more case (see INTEGRATION.md).
//
// This file is SOURCE ONLY in this repository: the build image has no Go toolchain.  It binds
// include/ptb200.h one to one; nothing here computes pixels.
//
// Build (in the reference tree): copy this directory to internal/engine/cuda, put ptb200.h on the include
// path and libptb200.so on the library path:
//
//	CGO_CFLAGS="-I/path/to/ptb200/include" CGO_LDFLAGS="-L/path/to/ptb200/path_trace_golang_b200 -lptb200" go build ./...
package cuda

/*
#include <stdlib.h>
#include "ptb200.h"

// ptbGoProgress is the Go function exported from callback.go.  cgo forbids DEFINITIONS in the preamble of a file that
// uses //export (that preamble is emitted twice), so the export lives in its own file and this file only wraps its address.
void ptbGoProgress(void* user);
static ptb_progress_fn ptb_go_progress_ptr(void) { return (ptb_progress_fn)ptbGoProgress; }
*/
import "C"

import (
	"errors"
	"fmt"
	"image"
	"hash/fnv"
	"math"
	"runtime/cgo"
	"sync"
	"unsafe"

	"github.com/user/pathtracer/internal/scene"
)

// RenderConfig is a minimal copy of engine.RenderConfig to avoid an import cycle (same trick as gpu.go:226-232).
type RenderConfig struct {
	Width        int
	Height       int
	SamplesPerPx int
	MaxDepth     int
}

// Seed is the key of the counter RNG (the CPU path seeds from the clock, random.go:14-16).
var Seed uint32 = 1

// Device is the CUDA device ordinal of the single-device context.  Changing it (or Devices) between renders takes effect
// on the next Render: the context is rebuilt under the lock.
var Device = 0

// Devices > 1 renders every frame on devices 0..Devices-1 of the box (ptb_multi_*: sample ranges per device, reduced on
// device 0).  The progress callback then fires only once, when the frame is complete.
var Devices = 1

var (
	mu         sync.Mutex // one render at a time per context (the GL path serialises on its worker, gpu.go:266-297)
	ctx        *C.ptb_ctx
	multi      *C.ptb_multi
	ctxDevice  = -1
	multiCount = 0
)

func lastError(c *C.ptb_ctx) error { return errors.New(C.GoString(C.ptb_last_error(c))) }

// ensureContext (called with mu held) creates the context for the current Device / Devices setting, replacing one made for
// other settings.  A failure is returned to the caller and retried on the next render (a GPU may have come back); nothing
// is rendered on the CPU instead.
func ensureContext() error {
	if Devices > 1 {
		if multi != nil && multiCount == Devices {
			return nil
		}
		if multi != nil {
			C.ptb_multi_destroy(multi)
			multi = nil
		}
		if rc := C.ptb_multi_create(nil, C.int(Devices), &multi); rc != C.PTB_OK {
			multi = nil
			return fmt.Errorf("CUDA initialization failed: %s", C.GoString(C.ptb_multi_last_error(nil)))
		}
		multiCount = Devices
		return nil
	}
	if ctx != nil && ctxDevice == Device {
		return nil
	}
	if ctx != nil {
		C.ptb_destroy(ctx)
		ctx = nil
	}
	if rc := C.ptb_create(C.int(Device), &ctx); rc != C.PTB_OK {
		ctx = nil
		return fmt.Errorf("CUDA initialization failed: %w", lastError(nil))
	}
	ctxDevice = Device
	return nil
}

func materialCode(t scene.MaterialType) C.int32_t { // materials.go:33-54
	switch t {
	case scene.MaterialMetal:
		return C.PTB_MAT_METAL
	case scene.MaterialDielectric:
		return C.PTB_MAT_DIELECTRIC
	case scene.MaterialEmissive:
		return C.PTB_MAT_EMISSIVE
	case scene.MaterialMirror:
		return C.PTB_MAT_MIRROR
	default:
		return C.PTB_MAT_LAMBERT
	}
}

func objectCode(t scene.ObjectType) C.int32_t { // objects.go:237-266
	switch t {
	case scene.ObjectSphere, scene.ObjectSphereLight:
		return C.PTB_OBJ_SPHERE
	case scene.ObjectPlane:
		return C.PTB_OBJ_PLANE
	case scene.ObjectBox:
		return C.PTB_OBJ_BOX
	case scene.ObjectMesh: // extension, go/internal/scene/mesh.go
		return C.PTB_OBJ_MESH
	default:
		return -1 // dropped by the library, like sceneToWorld drops unknown types
	}
}

// flatScene owns the C arrays of one ptb_scene.
type flatScene struct {
	s    C.ptb_scene
	free []unsafe.Pointer
}

func (f *flatScene) ints(v []int32) *C.int32_t {
	if len(v) == 0 {
		return nil
	}
	p := C.malloc(C.size_t(len(v) * 4))
	copy(unsafe.Slice((*int32)(p), len(v)), v)
	f.free = append(f.free, p)
	return (*C.int32_t)(p)
}

func (f *flatScene) doubles(v []float64) *C.double {
	if len(v) == 0 {
		return nil
	}
	p := C.malloc(C.size_t(len(v) * 8))
	copy(unsafe.Slice((*float64)(p), len(v)), v)
	f.free = append(f.free, p)
	return (*C.double)(p)
}

func (f *flatScene) floats(v []float32) *C.float {
	if len(v) == 0 {
		return nil
	}
	p := C.malloc(C.size_t(len(v) * 4))
	copy(unsafe.Slice((*float32)(p), len(v)), v)
	f.free = append(f.free, p)
	return (*C.float)(p)
}

func (f *flatScene) int64s(v []int64) *C.int64_t {
	if len(v) == 0 {
		return nil
	}
	p := C.malloc(C.size_t(len(v) * 8))
	copy(unsafe.Slice((*int64)(p), len(v)), v)
	f.free = append(f.free, p)
	return (*C.int64_t)(p)
}

func (f *flatScene) release() {
	for _, p := range f.free {
		C.free(p)
	}
}

// flatten turns a scene.Scene into the SoA view of ptb_scene: RAW fields only — convertMaterial, the box min/max
// arithmetic and newCamera run inside the library (include/ptb200.h).
func flatten(sc *scene.Scene) (*flatScene, error) {
	f := &flatScene{}
	byID := make(map[string]int32, len(sc.Materials)) // later duplicate wins (objects.go:226-229)
	nm := len(sc.Materials)
	mt := make([]int32, nm)
	alb, emit, abs := make([]float64, 3*nm), make([]float64, 3*nm), make([]float64, 3*nm)
	rough, ior, power, smooth := make([]float64, nm), make([]float64, nm), make([]float64, nm), make([]float64, nm)
	for i, m := range sc.Materials {
		byID[m.ID] = int32(i)
		mt[i] = int32(materialCode(m.Type))
		alb[3*i], alb[3*i+1], alb[3*i+2] = m.Albedo.R, m.Albedo.G, m.Albedo.B
		emit[3*i], emit[3*i+1], emit[3*i+2] = m.Emit.R, m.Emit.G, m.Emit.B
		abs[3*i], abs[3*i+1], abs[3*i+2] = m.Absorption.R, m.Absorption.G, m.Absorption.B
		rough[i], ior[i], power[i], smooth[i] = m.Rough, m.IOR, m.Power, m.Smoothness
	}
	no := len(sc.Objects)
	ot, om := make([]int32, no), make([]int32, no)
	pos, size := make([]float64, 3*no), make([]float64, 3*no)
	// meshes (extension): one entry per mesh object; triangles in world space, binary32 (ptb_scene.tri_vertices).
	// mesh_generation = hash of every mesh's (Generation, position, size): unchanged meshes cost nothing per frame.
	objMesh := make([]int32, no)
	triBegin := []int64{0}
	var tris []float32
	gen := fnv.New64a()
	for i := range sc.Objects {
		o := &sc.Objects[i]
		objMesh[i] = -1
		if o.Type != scene.ObjectMesh || o.Mesh == nil {
			continue
		}
		var err error
		if tris, err = o.WorldTriangles(tris); err != nil {
			f.release()
			return nil, err
		}
		if int64(len(tris)/9) == triBegin[len(triBegin)-1] {
			continue // no triangles: the object is dropped (its type code stays, the library skips empty meshes)
		}
		objMesh[i] = int32(len(triBegin) - 1)
		triBegin = append(triBegin, int64(len(tris)/9))
		var b [8 * 7]byte
		for k, v := range []uint64{o.Mesh.Generation(), math.Float64bits(o.Position.X), math.Float64bits(o.Position.Y), math.Float64bits(o.Position.Z),
			math.Float64bits(o.Size.X), math.Float64bits(o.Size.Y), math.Float64bits(o.Size.Z)} {
			for j := 0; j < 8; j++ {
				b[8*k+j] = byte(v >> (8 * j))
			}
		}
		gen.Write(b[:])
	}
	for i, o := range sc.Objects {
		ot[i] = int32(objectCode(o.Type))
		if o.Type == scene.ObjectMesh && objMesh[i] < 0 {
			ot[i] = -1
		}
		if idx, ok := byID[o.MaterialID]; ok {
			om[i] = idx
		} else {
			om[i] = -1 // zero material, objects.go:234
		}
		pos[3*i], pos[3*i+1], pos[3*i+2] = o.Position.X, o.Position.Y, o.Position.Z
		size[3*i], size[3*i+1], size[3*i+2] = o.Size.X, o.Size.Y, o.Size.Z
	}
	s := &f.s
	s.n_obj, s.obj_type, s.obj_mat, s.obj_pos, s.obj_size = C.int32_t(no), f.ints(ot), f.ints(om), f.doubles(pos), f.doubles(size)
	s.n_mat, s.mat_type = C.int32_t(nm), f.ints(mt)
	s.mat_albedo, s.mat_rough, s.mat_ior, s.mat_emit = f.doubles(alb), f.doubles(rough), f.doubles(ior), f.doubles(emit)
	s.mat_power, s.mat_absorption, s.mat_smoothness = f.doubles(power), f.doubles(abs), f.doubles(smooth)
	c := sc.Camera
	s.camera.position = [3]C.double{C.double(c.Position.X), C.double(c.Position.Y), C.double(c.Position.Z)}
	s.camera.target = [3]C.double{C.double(c.Target.X), C.double(c.Target.Y), C.double(c.Target.Z)}
	s.camera.up = [3]C.double{C.double(c.Up.X), C.double(c.Up.Y), C.double(c.Up.Z)}
	s.camera.fov, s.camera.aperture = C.double(c.FOV), C.double(c.Aperture)
	s.camera.focus_dist, s.camera.aspect_ratio = C.double(c.FocusDist), C.double(c.AspectRatio)
	if n := len(triBegin) - 1; n > 0 {
		s.n_mesh, s.obj_mesh, s.mesh_tri_begin, s.tri_vertices = C.int32_t(n), f.ints(objMesh), f.int64s(triBegin), f.floats(tris)
		s.mesh_generation = C.uint64_t(gen.Sum64() | 1)
	}
	// sky selection of renderIntoCPU, renderer.go:56-92
	if sc.Sky != nil && sc.Sky.Type == "gradient" {
		s.sky.kind = C.PTB_SKY_GRADIENT
		s.sky.horizon = [3]C.double{C.double(sc.Sky.Horizon.R), C.double(sc.Sky.Horizon.G), C.double(sc.Sky.Horizon.B)}
		s.sky.zenith = [3]C.double{C.double(sc.Sky.Zenith.R), C.double(sc.Sky.Zenith.G), C.double(sc.Sky.Zenith.B)}
	} else {
		bg := sc.Background
		if sc.Sky != nil && sc.Sky.Type == "solid" {
			bg = sc.Sky.Color
		}
		s.sky.kind = C.PTB_SKY_CONST
		s.sky.color = [3]C.double{C.double(bg.R), C.double(bg.G), C.double(bg.B)}
	}
	return f, nil
}

// Render renders sc into img on the CUDA backend.  Same contract as gpu.Render (gpu.go:2534-2546): the caller owns
// img; a size mismatch is a silent no-op like renderIntoCPU (renderer.go:46-49); progress may be nil.
// There is no CPU fallback: the error is returned to the caller.
func Render(sc *scene.Scene, cfg RenderConfig, img *image.RGBA, progress func()) error {
	b := img.Bounds()
	if b.Dx() != cfg.Width || b.Dy() != cfg.Height {
		return nil
	}
	mu.Lock()
	defer mu.Unlock()
	if err := ensureContext(); err != nil {
		return err
	}

	f, err := flatten(sc)
	if err != nil {
		return err
	}
	defer f.release()
	c := C.ptb_cfg{width: C.int32_t(cfg.Width), height: C.int32_t(cfg.Height), samples_per_px: C.int32_t(cfg.SamplesPerPx),
		max_depth: C.int32_t(cfg.MaxDepth), seed: C.uint32_t(Seed)}
	if multi != nil {
		if rc := C.ptb_multi_scene_upload(multi, &f.s); rc != C.PTB_OK {
			return fmt.Errorf("ptb_multi_scene_upload: %s", C.GoString(C.ptb_multi_last_error(multi)))
		}
		off := img.PixOffset(b.Min.X, b.Min.Y)
		if rc := C.ptb_multi_render(multi, &c, (*C.uint8_t)(unsafe.Pointer(&img.Pix[off])), C.size_t(img.Stride)); rc != C.PTB_OK {
			return fmt.Errorf("ptb_multi_render: %s", C.GoString(C.ptb_multi_last_error(multi)))
		}
		if progress != nil {
			progress()
		}
		return nil
	}
	if rc := C.ptb_scene_upload(ctx, &f.s); rc != C.PTB_OK {
		return fmt.Errorf("ptb_scene_upload: %w", lastError(ctx))
	}
	var cb C.ptb_progress_fn
	var user unsafe.Pointer
	if progress != nil {
		h := cgo.NewHandle(progress)
		defer h.Delete()
		cb, user = C.ptb_go_progress_ptr(), unsafe.Pointer(&h)
	}
	off := img.PixOffset(b.Min.X, b.Min.Y)
	if rc := C.ptb_render(ctx, &c, (*C.uint8_t)(unsafe.Pointer(&img.Pix[off])), C.size_t(img.Stride), cb, user); rc != C.PTB_OK {
		return fmt.Errorf("ptb_render: %w", lastError(ctx))
	}
	return nil
}
