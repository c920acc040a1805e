// gengolden — writes reference-produced golden vectors for the five shipped scenes (new command for the reference tree:
// copy to cmd/gengolden, with go/internal/engine/golden.go copied to internal/engine/).  SOURCE ONLY here: no Go toolchain in
// the build image.  Usage, from the root of the reference tree:
//
//	go run ./cmd/gengolden -out /path/to/ptb200/tests/golden/ref [-spp 4096]
//
// Output (little-endian, read by tests/test_reference_golden.py; see its docstring for the exact layout):
//
//	<scene>.primary.bin    int32[1080*1920] world index of the primary hit at the pixel centre (xi = 0.5, lens off), then
//	                       float64[1080*1920] its t
//	<scene>.camera.bin     float64[22] newCamera at 1920x1080 (pins math.Tan)
//	<scene>.converged.bin  float64[67*120*3] 4x4-block means (rows 0..267) of the 480x270 mean linear radiance at -spp samples
//	manifest.json          Go version, GOARCH, spp, seed, depths, world sizes
package main

import (
	"encoding/binary"
	"encoding/json"
	"flag"
	"log"
	"os"
	"path/filepath"
	"runtime"

	"github.com/user/pathtracer/internal/engine"
	"github.com/user/pathtracer/internal/scene"
)

var scenes = []struct {
	name  string
	depth int
}{{"example_simple", 8}, {"test_scene", 10}, {"metal_glass_room", 16}, {"test_comprehensive", 10}, {"gpu_showcase", 12}}

func writeBin(path string, parts ...interface{}) {
	f, err := os.Create(path)
	if err != nil {
		log.Fatal(err)
	}
	defer f.Close()
	for _, p := range parts {
		if err := binary.Write(f, binary.LittleEndian, p); err != nil {
			log.Fatal(err)
		}
	}
}

func main() {
	out := flag.String("out", "golden_ref", "output directory")
	sceneDir := flag.String("scenes", "scenes", "directory of the scene JSON files")
	spp := flag.Int("spp", 4096, "samples per pixel of the converged 480x270 render")
	seed := flag.Int64("seed", 777, "RNG seed of the converged render (row y uses seed+y)")
	flag.Parse()
	if err := os.MkdirAll(*out, 0o755); err != nil {
		log.Fatal(err)
	}
	manifest := map[string]interface{}{"go": runtime.Version(), "goarch": runtime.GOARCH, "spp": *spp, "seed": *seed,
		"primary": map[string]interface{}{"width": 1920, "height": 1080, "xi_u": 0.5, "xi_v": 0.5},
		"converged": map[string]interface{}{"width": 480, "height": 270, "block": 4, "rows_used": 268}}
	worlds := map[string]int{}
	depths := map[string]int{}
	for _, s := range scenes {
		sc, err := scene.Load(filepath.Join(*sceneDir, s.name+".json"))
		if err != nil {
			log.Fatal(err)
		}
		full := engine.RenderConfig{Width: 1920, Height: 1080, SamplesPerPx: 1, MaxDepth: s.depth}
		ids, ts := engine.GoldenPrimaryHits(sc, full, 0.5, 0.5)
		writeBin(filepath.Join(*out, s.name+".primary.bin"), ids, ts)
		cam := engine.GoldenCamera(sc, full)
		writeBin(filepath.Join(*out, s.name+".camera.bin"), cam[:])
		small := engine.RenderConfig{Width: 480, Height: 270, SamplesPerPx: *spp, MaxDepth: s.depth}
		lin := engine.GoldenRenderLinear(sc, small, *seed)
		blocks := make([]float64, 67*120*3)
		for by := 0; by < 67; by++ {
			for bx := 0; bx < 120; bx++ {
				for c := 0; c < 3; c++ {
					sum := 0.0
					for dy := 0; dy < 4; dy++ {
						for dx := 0; dx < 4; dx++ {
							sum += lin[((by*4+dy)*480+bx*4+dx)*3+c]
						}
					}
					blocks[(by*120+bx)*3+c] = sum / 16
				}
			}
		}
		writeBin(filepath.Join(*out, s.name+".converged.bin"), blocks)
		worlds[s.name] = engine.GoldenWorldSize(sc)
		depths[s.name] = s.depth
		log.Printf("%s: done", s.name)
	}
	manifest["world_size"] = worlds
	manifest["max_depth"] = depths
	b, _ := json.MarshalIndent(manifest, "", "  ")
	if err := os.WriteFile(filepath.Join(*out, "manifest.json"), b, 0o644); err != nil {
		log.Fatal(err)
	}
}
