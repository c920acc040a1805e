// mesh.go — triangle-mesh extension of the scene schema (new file for internal/scene; go/patches/scene.go.patch adds the
// Object.Mesh field and the ObjectMesh type string).  SOURCE ONLY in this repository: the build image has no Go toolchain.
// The C++ host mirror (path_trace_golang_b200/csrc/host/scene.cpp) implements the same schema and the same generator —
// tests/test_host_scene.py pins that one; this file restates it for the reference tree.
//
// JSON (backward compatible: encoding/json ignores unknown keys in older binaries, sceneToWorld drops the unknown type):
//
//	{"type": "mesh", "position": {...}, "size": {...}, "material_id": "...",
//	 "mesh": {"vertices": [x0,y0,z0, ...], "triangles": [a0,b0,c0, ...]}}                      inline, or
//	 "mesh": {"heightfield": {"nx":1000,"nz":500,"seed":1234,"amplitude":0.4,"frequency":3,"octaves":4}}
//
// World vertex = position + size ⊙ local vertex (a zero size component means 1).
package scene

import (
	"errors"
	"math"
	"sync/atomic"
)

// Heightfield parameters of the deterministic terrain generator (the C4 benchmark workload).
type Heightfield struct {
	NX        int     `json:"nx"`
	NZ        int     `json:"nz"`
	Seed      uint32  `json:"seed"`
	Amplitude float64 `json:"amplitude"`
	Frequency float64 `json:"frequency"`
	Octaves   int     `json:"octaves"`
}

// Mesh is the geometry of an Object of type ObjectMesh.
type Mesh struct {
	Vertices    []float32    `json:"vertices,omitempty"`  // 3 per vertex, local space
	Triangles   []uint32     `json:"triangles,omitempty"` // 3 vertex indices per triangle
	Heightfield *Heightfield `json:"heightfield,omitempty"`

	generation uint64 // identity of Vertices/Triangles for ptb_scene.mesh_generation; see Touch
}

var nextGeneration uint64

// Generation identifies the current contents of the mesh: it changes only through Touch (or the first call).
func (m *Mesh) Generation() uint64 {
	if g := atomic.LoadUint64(&m.generation); g != 0 {
		return g
	}
	g := atomic.AddUint64(&nextGeneration, 1)
	if atomic.CompareAndSwapUint64(&m.generation, 0, g) {
		return g
	}
	return atomic.LoadUint64(&m.generation)
}

// Touch must be called after Vertices or Triangles were edited in place: the CUDA backend recognises an unchanged mesh by its
// generation and then neither re-reads the triangles nor rebuilds the BVH.
func (m *Mesh) Touch() { atomic.StoreUint64(&m.generation, atomic.AddUint64(&nextGeneration, 1)) }

func lattice(ix, iz int32, seed uint32, oct int) float64 {
	h := uint32(ix)*0x9E3779B1 ^ uint32(iz)*0x85EBCA77 ^ (seed + uint32(oct)*0xC2B2AE3D)
	h ^= h >> 16
	h *= 0x21f0aaad
	h ^= h >> 15
	h *= 0x735a2d97
	h ^= h >> 15
	return float64(h)*(2.0/4294967296.0) - 1.0
}

func vnoise(x, z float64, seed uint32, oct int) float64 {
	fx, fz := math.Floor(x), math.Floor(z)
	ix, iz := int32(fx), int32(fz)
	tx, tz := x-fx, z-fz
	tx = tx * tx * (3.0 - 2.0*tx)
	tz = tz * tz * (3.0 - 2.0*tz)
	a, b := lattice(ix, iz, seed, oct), lattice(ix+1, iz, seed, oct)
	c, d := lattice(ix, iz+1, seed, oct), lattice(ix+1, iz+1, seed, oct)
	return (a + (b-a)*tx) + ((c+(d-c)*tx)-(a+(b-a)*tx))*tz
}

// Generate fills Vertices/Triangles from Heightfield: nx × nz quads (two triangles each) over [-0.5, 0.5]² in x, z with
// y = amplitude · Σ_o 0.5^o · vnoise((x+0.5)·f·2^o, (z+0.5)·f·2^o); binary64 arithmetic, binary32 vertices.
func (m *Mesh) Generate() error {
	hf := m.Heightfield
	if hf == nil {
		return nil
	}
	if hf.NX < 1 || hf.NZ < 1 || int64(hf.NX)*int64(hf.NZ) > 50000000 {
		return errors.New("decode scene: heightfield nx/nz out of range")
	}
	oct := hf.Octaves
	if oct <= 0 {
		oct = 4
	}
	freq := hf.Frequency
	if freq == 0 {
		freq = 4.0
	}
	nx, nz := hf.NX, hf.NZ
	m.Vertices = make([]float32, (nx+1)*(nz+1)*3)
	for j := 0; j <= nz; j++ {
		for i := 0; i <= nx; i++ {
			x, z := float64(i)/float64(nx)-0.5, float64(j)/float64(nz)-0.5
			y, w, f := 0.0, 1.0, freq
			for o := 0; o < oct; o++ {
				y += w * vnoise((x+0.5)*f, (z+0.5)*f, hf.Seed, o)
				w *= 0.5
				f *= 2
			}
			k := (j*(nx+1) + i) * 3
			m.Vertices[k], m.Vertices[k+1], m.Vertices[k+2] = float32(x), float32(hf.Amplitude*y), float32(z)
		}
	}
	m.Triangles = make([]uint32, nx*nz*6)
	k := 0
	for j := 0; j < nz; j++ {
		for i := 0; i < nx; i++ {
			a := uint32(j*(nx+1) + i)
			b, c := a+1, a+uint32(nx+1)
			d := c + 1
			m.Triangles[k], m.Triangles[k+1], m.Triangles[k+2] = a, c, b // counter-clockwise seen from +y
			m.Triangles[k+3], m.Triangles[k+4], m.Triangles[k+5] = b, c, d
			k += 6
		}
	}
	m.Touch()
	return nil
}

// WorldTriangles appends the object's triangles in world space, binary32, 9 floats each (v0, v1, v2) — the layout of
// ptb_scene.tri_vertices.  Generates a heightfield on first use.
func (o *Object) WorldTriangles(dst []float32) ([]float32, error) {
	m := o.Mesh
	if o.Type != ObjectMesh || m == nil {
		return dst, nil
	}
	if len(m.Triangles) == 0 && m.Heightfield != nil {
		if err := m.Generate(); err != nil {
			return dst, err
		}
	}
	sx, sy, sz := o.Size.X, o.Size.Y, o.Size.Z
	if sx == 0 {
		sx = 1
	}
	if sy == 0 {
		sy = 1
	}
	if sz == 0 {
		sz = 1
	}
	for _, idx := range m.Triangles {
		if int(idx)*3+2 >= len(m.Vertices) {
			return dst, errors.New("mesh: triangle index out of range")
		}
		dst = append(dst,
			float32(o.Position.X+sx*float64(m.Vertices[3*idx])),
			float32(o.Position.Y+sy*float64(m.Vertices[3*idx+1])),
			float32(o.Position.Z+sz*float64(m.Vertices[3*idx+2])))
	}
	return dst, nil
}
