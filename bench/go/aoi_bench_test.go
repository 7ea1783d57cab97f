// UN-RUN (no Go toolchain in the build image).  Drop into pkg/channeld of channeldorg/channeld @ 61fa8add and run
//   go test ./pkg/channeld -run XXX -bench BenchmarkSphereAOIQueries -benchtime 10x
// to obtain the genuine Go figure for BASELINE.json's metric on the synthetic world of SURVEY.md §8d
// (config #2: spatial_static_benchmark.json, 1 M entities, 100 K subscribers, r = 50).
package channeld

import (
	"testing"

	"github.com/channeldorg/channeld/pkg/channeldpb"
	"github.com/channeldorg/channeld/pkg/common"
)

func splitmix64(x uint64) uint64 {
	z := x + 0x9E3779B97F4A7C15
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

func uniform(seed, i, k uint64) float64 {
	return float64(splitmix64(seed*0x9E3779B97F4A7C15+2*i+k)>>11) * (1.0 / 9007199254740992.0)
}

func BenchmarkSphereAOIQueries(b *testing.B) {
	ctl := &StaticGrid2DSpatialController{WorldOffsetX: -15000, WorldOffsetZ: -15000, GridWidth: 2000, GridHeight: 2000,
		GridCols: 15, GridRows: 15, ServerCols: 3, ServerRows: 3}
	const N, S, seed = 1000000, 100000, 2
	ex, ez := make([]float64, N), make([]float64, N)
	cells := make(map[common.ChannelId][]uint32)
	for i := 0; i < N; i++ {
		ex[i] = -15000 + uniform(seed, uint64(i), 0)*30000
		ez[i] = -15000 + uniform(seed, uint64(i), 1)*30000
		if id, err := ctl.GetChannelId(common.SpatialInfo{X: ex[i], Z: ez[i]}); err == nil {
			cells[id] = append(cells[id], uint32(i))
		}
	}
	b.ResetTimer()
	b.RunParallel(func(pb *testing.PB) {
		j := 0
		var visible []uint32
		for pb.Next() {
			e := (j % S) * (N / S)
			q := &channeldpb.SpatialInterestQuery{SphereAOI: &channeldpb.SpatialInterestQuery_SphereAOI{
				Center: &channeldpb.SpatialInfo{X: ex[e], Z: ez[e]}, Radius: 50}}
			res, _ := ctl.QueryChannelIds(q)
			visible = visible[:0]
			for id := range res {
				visible = append(visible, cells[id]...)
			}
			j++
		}
	})
}
