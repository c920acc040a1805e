// go_cpu_bench_test.go — times the reference's own CPU renderer (engine.Render with BackendCPU) on the BASELINE.md
// configurations, for anyone with a Go >= 1.21 toolchain (this repository's build image has none, so bench.py
// reports a C++ restatement instead and labels it "port").
//
// Drop into the reference tree as internal/engine/go_cpu_bench_test.go and run
//
//	go test ./internal/engine -run XXX -bench CPU -benchtime 1x
//
// Reports Msamples/s; workers = runtime.NumCPU() (override with PATHTRACER_WORKERS, renderer.go:123-129).
package engine

import (
	"runtime"
	"testing"

	"github.com/user/pathtracer/internal/scene"
)

func benchCPU(b *testing.B, path string, w, h, spp, depth int) {
	sc, err := scene.Load(path)
	if err != nil {
		b.Fatal(err)
	}
	SetBackend(BackendCPU)
	cfg := RenderConfig{Width: w, Height: h, SamplesPerPx: spp, MaxDepth: depth}
	b.ResetTimer()
	for i := 0; i < b.N; i++ {
		Render(sc, cfg)
	}
	samples := float64(w) * float64(h) * float64(spp) * float64(b.N)
	b.ReportMetric(samples/b.Elapsed().Seconds()/1e6, "Msamples/s")
	b.ReportMetric(float64(runtime.NumCPU()), "cores")
}

func BenchmarkCPU_C1(b *testing.B) { benchCPU(b, "../../scenes/example_simple.json", 640, 360, 16, 8) }
func BenchmarkCPU_C2(b *testing.B) { benchCPU(b, "../../scenes/test_scene.json", 1920, 1080, 64, 10) }

// C3 at 4 of its 256 spp: samples/s does not depend on spp.
func BenchmarkCPU_C3(b *testing.B) { benchCPU(b, "../../scenes/metal_glass_room.json", 3840, 2160, 4, 16) }
