// golden.go — golden-vector exports for pinning ports of the CPU renderer (new file for internal/engine).
// SOURCE ONLY in this repository: the build image has no Go toolchain, so these functions have never been compiled here.
// They exist so that the first box WITH a Go toolchain can turn the "parity unpinned" status of ../../oracle into "pinned":
// go/cmd/gengolden writes what they return into tests/golden/ref/, and tests/test_reference_golden.py compares both the
// C++ oracle and the CUDA device against those files.
//
// Everything here calls the reference's own unexported code (sceneToWorld, newCamera, camera.getRay, hittable.hit,
// rayColorOpt): nothing is re-derived.
package engine

import (
	"math"
	"math/rand"
	"runtime"
	"sync"

	"github.com/user/pathtracer/internal/scene"
)

// GoldenCamera returns newCamera's result for the frame size: origin, lowerLeftCorner, horizontal, vertical, u, v, w
// (3 float64 each) and lensRadius — 22 values.  It pins math.Tan (camera.go:26), the one libm call in the geometry.
func GoldenCamera(sc *scene.Scene, cfg RenderConfig) [22]float64 {
	c := newCamera(sc.Camera, cfg, nil)
	var out [22]float64
	for i, p := range []vec3{c.origin, c.lowerLeftCorner, c.horizontal, c.vertical, c.u, c.v, c.w} {
		out[3*i], out[3*i+1], out[3*i+2] = p.x, p.y, p.z
	}
	out[21] = c.lensRadius
	return out
}

// GoldenPrimaryHits returns, for every pixel (row-major, top row first), the index in sceneToWorld's slice of the object
// the primary ray through (x+xiU, y+xiV) hits first (-1: none) and its ray parameter t (0 on a miss).  The camera is built
// with a nil rng, i.e. getRay's lens-free branch (camera.go:70-73); the scan is the one of rayColorOpt (renderer.go:292-302):
// tMin = 0.001, closest starts at MaxFloat64, every object tested in order, a hit replaces the record.
func GoldenPrimaryHits(sc *scene.Scene, cfg RenderConfig, xiU, xiV float64) (ids []int32, ts []float64) {
	world := sceneToWorld(sc)
	cam := newCamera(sc.Camera, cfg, nil)
	invWidth := 1.0 / float64(cfg.Width-1)
	invHeight := 1.0 / float64(cfg.Height-1)
	heightMinus1 := float64(cfg.Height - 1)
	ids = make([]int32, cfg.Width*cfg.Height)
	ts = make([]float64, cfg.Width*cfg.Height)
	var wg sync.WaitGroup
	rows := make(chan int, cfg.Height)
	for y := 0; y < cfg.Height; y++ {
		rows <- y
	}
	close(rows)
	for w := 0; w < runtime.NumCPU(); w++ {
		wg.Add(1)
		go func() {
			defer wg.Done()
			var rec hitRecord
			for y := range rows {
				flipY := heightMinus1 - float64(y)
				for x := 0; x < cfg.Width; x++ {
					u := (float64(x) + xiU) * invWidth
					vv := (flipY + xiV) * invHeight
					r := cam.getRay(u, vv)
					const tMin = 0.001
					closest := math.MaxFloat64
					id := int32(-1)
					for i := range world {
						if world[i].hit(r, tMin, closest, &rec) {
							closest = rec.t
							id = int32(i)
						}
					}
					ids[y*cfg.Width+x] = id
					if id >= 0 {
						ts[y*cfg.Width+x] = closest
					}
				}
			}
		}()
	}
	wg.Wait()
	return ids, ts
}

// goldenBackground is the sky selection of renderIntoCPU (renderer.go:56-92), which lives in a closure there.
func goldenBackground(sc *scene.Scene) func(ray) vec3 {
	if sc.Sky != nil && sc.Sky.Type == "gradient" {
		horizon := v(sc.Sky.Horizon.R, sc.Sky.Horizon.G, sc.Sky.Horizon.B)
		zenith := v(sc.Sky.Zenith.R, sc.Sky.Zenith.G, sc.Sky.Zenith.B)
		return func(r ray) vec3 {
			dirLen := math.Sqrt(r.dir.x*r.dir.x + r.dir.y*r.dir.y + r.dir.z*r.dir.z)
			if dirLen == 0 {
				return horizon
			}
			t := (r.dir.y/dirLen + 1.0) * 0.5
			if t < 0 {
				t = 0
			}
			if t > 1 {
				t = 1
			}
			return vec3{x: horizon.x*(1-t) + zenith.x*t, y: horizon.y*(1-t) + zenith.y*t, z: horizon.z*(1-t) + zenith.z*t}
		}
	}
	bg := v(sc.Background.R, sc.Background.G, sc.Background.B)
	if sc.Sky != nil && sc.Sky.Type == "solid" {
		bg = v(sc.Sky.Color.R, sc.Sky.Color.G, sc.Sky.Color.B)
	}
	return func(r ray) vec3 { return bg }
}

// GoldenRenderLinear renders the per-pixel MEAN linear radiance (the `col` of renderer.go:176-192 after the division by the
// sample count, before sqrt and quantisation): width*height*3 float64, row-major, top row first.  The sample loop is the
// reference's (u-jitter, v-jitter, getRay with the lens, rayColorOpt); only the scheduling differs: one rand.Rand per ROW
// seeded with seed+y, so the result is a pure function of (scene, cfg, seed) whatever the number of cores.
func GoldenRenderLinear(sc *scene.Scene, cfg RenderConfig, seed int64) []float64 {
	world := sceneToWorld(sc)
	bg := goldenBackground(sc)
	invWidth := 1.0 / float64(cfg.Width-1)
	invHeight := 1.0 / float64(cfg.Height-1)
	invSamples := 1.0 / float64(cfg.SamplesPerPx)
	heightMinus1 := float64(cfg.Height - 1)
	out := make([]float64, cfg.Width*cfg.Height*3)
	var wg sync.WaitGroup
	rows := make(chan int, cfg.Height)
	for y := 0; y < cfg.Height; y++ {
		rows <- y
	}
	close(rows)
	for w := 0; w < runtime.NumCPU(); w++ {
		wg.Add(1)
		go func() {
			defer wg.Done()
			for y := range rows {
				rng := &randSource{r: rand.New(rand.NewSource(seed + int64(y)))}
				cam := newCamera(sc.Camera, cfg, rng)
				flipY := heightMinus1 - float64(y)
				for x := 0; x < cfg.Width; x++ {
					col := vec3{}
					xFloat := float64(x)
					for s := 0; s < cfg.SamplesPerPx; s++ {
						u := (xFloat + rng.Float64()) * invWidth
						vv := (flipY + rng.Float64()) * invHeight
						r := cam.getRay(u, vv)
						var rec hitRecord
						col = col.add(rayColorOpt(r, world, bg, cfg.MaxDepth, rng, &rec))
					}
					k := (y*cfg.Width + x) * 3
					out[k], out[k+1], out[k+2] = col.x*invSamples, col.y*invSamples, col.z*invSamples
				}
			}
		}()
	}
	wg.Wait()
	return out
}

// GoldenWorldSize is len(sceneToWorld(sc)): the number of kept objects (unknown types are dropped, objects.go:237-266).
func GoldenWorldSize(sc *scene.Scene) int { return len(sceneToWorld(sc)) }
