// kernels_fused_tma.cuh -- the one-pass timestep kernel with TMA-staged inputs.
//
// Same tiling, arithmetic and results as k_fused_EH (kernels_fused.cuh): block = 128 x cells by
// FUSED_TY rows (+ one halo-row warp), marching a z chunk upwards, E_new and H_new of every cell
// produced in one pass.  The difference is how the inputs arrive: instead of every thread
// loading its own float4s into registers and waiting for them, ONE thread issues three bulk
// tensor copies per plane (TMA, cp.async.bulk.tensor) that land the tile of H_old (3 x 9 rows),
// E_old (3 x 8 rows) and the operator index (8 rows), 136 values wide (x0-4 .. x0+131: the x-1
// neighbour and the halo column x0+128 included), in a ring of shared-memory stages, signalled
// by an mbarrier per stage.  The copy of plane kk+1 (.. kk+STAGES-1) is in flight while plane kk
// is computed, so HBM latency is covered by the ring (30 KB per block and stage) and not by
// occupancy; no thread holds load results in registers across a wait.
// Out-of-range rows / columns (x0-4 < 0, j0-1 < 0, beyond the pitch or ny) are zero-filled by
// the TMA unit; the clamped neighbours of index 0 (engine.cpp:117,122,127) are taken from the
// own cell as before.
#pragma once
#include <cuda.h>
#include "kernels_fused.cuh"

#ifndef FT_STAGES
#define FT_STAGES 2                   // ring depth; 2 measured faster than 3 (more of the SM's 256 KB left as L1)
#endif
#ifndef FT_MIN_BLOCKS
#define FT_MIN_BLOCKS 2
#endif
#define FT_W 136                      // staged row: x0-4 .. x0+131
#define FT_ROWS_I (FUSED_TY + 2)      // H rows j0-1 .. j0+TY
#define FT_ROWS_V (FUSED_TY + 1)      // E / index rows j0 .. j0+TY

// UPML box that is thin in x and sits at the low / high end of the mesh ("x slab"), updated by its own
// one-pass kernel k_xslab_EH (kernels_xslab.cuh) instead of the shell launches.  For the big kernel
// these are cells it must not store: fp* = the footprint (whole float4 chunks, rows, local planes).
struct XSlabFoot {
	int on;
	int c0, cn;          // float4 chunks of the footprint
	int j0, jn, k0, kn;  // rows / local planes of the box
	int x0, x1;          // lines of the box itself [x0, x1)
};

// Lorentz / Drude ADE applied inside the one-pass kernel (template flag LOR): per order the auxiliary values the list
// kernels k_lorentz_pre have just advanced (engine_ext_lorentzmaterial.cpp:79-127), and per (plane, row) the one x
// segment of dispersive cells: rows[k * ny + j] = {x0 | n << 16, list index of the first}.  E_new = stencil - ADE
// (Apply2Voltages, :129-148) before H is computed from it, H_new = stencil - ADE (Apply2Current, :150-168).
#define LOR_FUSED_MAX 2
struct LorFusedOrder {
	const int2* rows;
	const float* ade_v;   // [3][count] or NULL
	const float* ade_i;
	unsigned count;
};
struct LorFusedParams {
	int nord;
	LorFusedOrder o[LOR_FUSED_MAX];
};

struct alignas(64) FusedTmaParams {
	LorFusedParams lor;
	XSlabFoot xs[2];
	CUtensorMap mI;   // H_old of the source set: (x, y, z, component) float, box 136 x 9 x 1 x 3
	CUtensorMap mV;   // E_old of the source set (shell cells: already E_new), box 136 x 8 x 1 x 3
	CUtensorMap mX;   // operator index (x, y, z), box 136 x 8 x 1
	FusedParams f;
};

template <typename IdxT> struct FtStage {
	static constexpr int I_BYTES = 3 * FT_ROWS_I * FT_W * 4;
	static constexpr int V_BYTES = 3 * FT_ROWS_V * FT_W * 4;
	static constexpr int X_BYTES = FT_ROWS_V * FT_W * (int)sizeof(IdxT);
	static constexpr int V_OFF = (I_BYTES + 127) / 128 * 128;
	static constexpr int X_OFF = V_OFF + (V_BYTES + 127) / 128 * 128;
	static constexpr int BYTES = X_OFF + (X_BYTES + 127) / 128 * 128;
	static constexpr int TX = I_BYTES + V_BYTES + X_BYTES; // bytes one stage receives
};
template <typename IdxT, int STAGES> constexpr int ft_smem_bytes()
{
	return STAGES * FtStage<IdxT>::BYTES + 2 * 3 * (FUSED_TY + 1) * 32 * 16 + 64;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "FT_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	    "@P1 bra FT_DONE;\n"
	    "bra FT_WAIT;\n"
	    "FT_DONE:\n"
	    "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <typename IdxT> struct SIdx4;
template <> struct SIdx4<uint16_t> {
	__device__ __forceinline__ static void load(const unsigned char* row, int lane, unsigned e[4])
	{
		const ushort4 v = reinterpret_cast<const ushort4*>(row)[lane];
		e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
	}
};
template <> struct SIdx4<uint32_t> {
	__device__ __forceinline__ static void load(const unsigned char* row, int lane, unsigned e[4])
	{
		const uint4 v = reinterpret_cast<const uint4*>(row)[lane];
		e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
	}
};

template <typename IdxT, bool HAS_PML, int STAGES, bool LOR = false>
__global__ void __launch_bounds__(32 * (FUSED_TY + 1), FT_MIN_BLOCKS) k_fused_tma(const __grid_constant__ FusedTmaParams P)
{
	PDL_PROLOGUE();
	extern __shared__ __align__(128) unsigned char ft_smem[];
	typedef FtStage<IdxT> ST;
	const FusedParams& p = P.f;
	float4 (*xV0)[FUSED_TY + 1][32] = reinterpret_cast<float4 (*)[FUSED_TY + 1][32]>(ft_smem + STAGES * ST::BYTES);
	float4 (*xV2)[FUSED_TY + 1][32] = xV0 + 3;
	const uint32_t bar0 = smem_u32(ft_smem + STAGES * ST::BYTES + 2 * 3 * (FUSED_TY + 1) * 32 * 16);

	const int lane = threadIdx.x, ty = threadIdx.y;
	const int x0 = blockIdx.x * 128, j0 = p.jb + blockIdx.y * FUSED_TY;
	const int i0 = x0 + lane * 4;
	const int j = j0 + ty;
	const bool halo_row = ty == FUSED_TY || j >= p.je;   // rows whose E is only computed for the row below them
	const bool producer = ty == FUSED_TY && lane == 0;
	const int kb = p.kE0 + blockIdx.z * p.zchunk;
	const int ke = min(kb + p.zchunk, p.kE1);
	if (kb >= ke) return; // block-uniform
	const int he = min(ke, p.kH1);
	const int e_last = (he == ke && ke < p.nz) ? ke : ke - 1;

	if (producer) {
		for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto issue = [&](int kk) {
		const int s = (kk - kb) % STAGES;
		const uint32_t bar = bar0 + 8 * s;
		const uint32_t dst = smem_u32(ft_smem + s * ST::BYTES);
		mbar_expect_tx(bar, ST::TX);
		tma_load_4d(dst, &P.mI, x0 - 4, j0 - 1, kk, 0, bar);
		tma_load_4d(dst + ST::V_OFF, &P.mV, x0 - 4, j0, kk, 0, bar);
		tma_load_3d(dst + ST::X_OFF, &P.mX, x0, j0, kk, bar);
	};
	if (producer)
		for (int s = 0; s < STAGES && kb + s <= e_last; ++s) issue(kb + s);

	const bool row_ok = j < p.ny;
	const bool active = row_ok && i0 < p.pitch;
	const int ic = i0;
	const long long row = (long long)(row_ok ? j : 0) * p.pitch + (i0 < p.pitch ? i0 : 0);
	// rows of the staged boxes: H box row r holds y = j0-1+r, E / index box row r holds y = j0+r
	const int rI = ty + 1, rIm = (j > 0) ? ty : ty + 1, rV = ty;
	const int xe = ic + 4;
	const bool hcol = lane == 31 && !halo_row && active && xe < p.nx;

	unsigned long long shb = 0, shxb = 0;
	if (HAS_PML) {
		const int chunk = ic >> 2;
		const int jc = row_ok ? j : p.ny - 1;
		for (int b = 0; b < p.nsh; ++b) {
			if ((unsigned)(jc - p.sh[b].j0) >= (unsigned)p.sh[b].jn) continue;
			const bool mine = (unsigned)(chunk - p.sh[b].c0) < (unsigned)p.sh[b].cn;
			const bool right = (unsigned)(chunk + 1 - p.sh[b].c0) < (unsigned)p.sh[b].cn;
			if (!mine && !right) continue;
			const int a = max(p.sh[b].k0, kb) - kb, z = min(p.sh[b].k0 + p.sh[b].kn, e_last + 1) - kb;
			if (z <= a) continue;
			const unsigned long long m = (z - a >= 64 ? ~0ull : ((1ull << (z - a)) - 1ull)) << a;
			if (mine) shb |= m;
			if (right) shxb |= m;
		}
	}
	// x-slab footprints: cells the x-slab kernel stores (one bit per plane of the march, like shb)
	unsigned long long xhb = 0;
	if (HAS_PML) {
		const int chunk = ic >> 2;
		const int jc = row_ok ? j : p.ny - 1;
		for (int g = 0; g < 2; ++g) {
			const XSlabFoot& X = P.xs[g];
			if (!X.on || (unsigned)(chunk - X.c0) >= (unsigned)X.cn || (unsigned)(jc - X.j0) >= (unsigned)X.jn) continue;
			const int a = max(X.k0, kb) - kb, z = min(X.k0 + X.kn, e_last + 1) - kb;
			if (z <= a) continue;
			xhb |= (z - a >= 64 ? ~0ull : ((1ull << (z - a)) - 1ull)) << a;
		}
	}
	bool shk = false, xsk = false;

	float4 ek0 = make_float4(0, 0, 0, 0), ek1 = ek0, ek2 = ek0;
	float4 hk0 = ek0, hk1 = ek0, hk2 = ek0;
	unsigned ek_idx[4] = {0, 0, 0, 0};
	float hcV1 = 0.0f, hcV2 = 0.0f, hcI0 = 0.0f;
	if (active || (halo_row && row_ok && i0 < p.pitch)) {
		const int km = kb - (kb > 0); // z-1 clamp only at the bottom of the (local) domain
		const long long o = (long long)km * p.plane + row;
		hk0 = ld4(p.Is + o);
		hk1 = ld4(p.Is + p.comp + o);
		if (hcol) hcI0 = p.Is[o + 4];
	}

	int2 lorR[LOR ? LOR_FUSED_MAX : 1], lorRk[LOR ? LOR_FUSED_MAX : 1]; // row segments of plane kk / plane kk-1
	if (LOR)
		for (int o = 0; o < LOR_FUSED_MAX; ++o) lorR[o] = lorRk[o] = make_int2(0, 0);
	const long long lor_row = row_ok ? j : p.ny - 1;

	for (int kk = kb; kk <= e_last; ++kk) {
		// ------------------------------------------------------------ E_new(kk)
		if (LOR) {
#pragma unroll
			for (int o = 0; o < LOR_FUSED_MAX; ++o)
				if (o < P.lor.nord) lorR[o] = __ldg(P.lor.o[o].rows + (long long)kk * p.ny + lor_row);
		}
		const int s = (kk - kb) % STAGES;
		mbar_wait(bar0 + 8 * s, ((kk - kb) / STAGES) & 1);
		const unsigned char* stg = ft_smem + s * ST::BYTES;
		const float* sI = reinterpret_cast<const float*>(stg);
		const float* sV = reinterpret_cast<const float*>(stg + ST::V_OFF);
		const unsigned char* sX = stg + ST::X_OFF;
		const bool sh = HAS_PML && (shb >> (kk - kb) & 1ull), shx = HAS_PML && (shxb >> (kk - kb) & 1ull);
		const bool xsc = HAS_PML && (xhb >> (kk - kb) & 1ull);
		const int cI = FT_ROWS_I * FT_W, cV = FT_ROWS_V * FT_W; // component strides inside a stage
		const int oI = rI * FT_W + 4 + lane * 4, oIm = rIm * FT_W + 4 + lane * 4, oV = rV * FT_W + 4 + lane * 4;
		unsigned e[4];
		SIdx4<IdxT>::load(sX + rV * FT_W * (int)sizeof(IdxT), lane, e);
		const float4 i0c = ld4(sI + oI), i1c = ld4(sI + cI + oI), i2c = ld4(sI + 2 * cI + oI);
		const float4 i0jm = ld4(sI + oIm), i2jm = ld4(sI + 2 * cI + oIm);
		float4 v0 = ld4(sV + oV), v1 = ld4(sV + cV + oV), v2 = ld4(sV + 2 * cV + oV); // shell cells: already E_new
		float l1 = __shfl_up_sync(0xffffffffu, i1c.w, 1);
		float l2 = __shfl_up_sync(0xffffffffu, i2c.w, 1);
		if (lane == 0) {
			if (ic > 0) { l1 = sI[cI + oI - 1]; l2 = sI[2 * cI + oI - 1]; }
			else { l1 = i1c.x; l2 = i2c.x; }
		}
		const float4 i1xm = make_float4(l1, i1c.x, i1c.y, i1c.z);
		const float4 i2xm = make_float4(l2, i2c.x, i2c.y, i2c.z);
		float nV1 = 0.0f, nV2 = 0.0f, nI0 = 0.0f; // halo column values of plane kk
		if (active) {
			const bool uni = (e[0] == e[1]) & (e[1] == e[2]) & (e[2] == e[3]);
			float4 A = __ldg(p.eA + e[0]), B = __ldg(p.eB + e[0]);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				if (c > 0 && !uni) { A = __ldg(p.eA + e[c]); B = __ldg(p.eB + e[c]); }
				const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(hk1, c));
				const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(hk0, c)), comp(i2c, c)), comp(i2xm, c));
				const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
				const float n0 = leap(comp(v0, c), A.x, B.x, curl0), n1 = leap(comp(v1, c), A.y, B.y, curl1), n2 = leap(comp(v2, c), A.z, B.z, curl2);
				setcomp(v0, c, sh ? comp(v0, c) : n0);
				setcomp(v1, c, sh ? comp(v1, c) : n1);
				setcomp(v2, c, sh ? comp(v2, c) : n2);
			}
			if (LOR) {
#pragma unroll
				for (int o = 0; o < LOR_FUSED_MAX; ++o) {
					const LorFusedOrder& Lo = P.lor.o[o];
					const int d = ic - (lorR[o].x & 0xffff), n = lorR[o].x >> 16;
					if (o < P.lor.nord && Lo.ade_v && d + 3 >= 0 && d < n) {
						const float* a = Lo.ade_v + lorR[o].y + d;
#pragma unroll
						for (int c = 0; c < 4; ++c)
							if ((unsigned)(d + c) < (unsigned)n) {
								setcomp(v0, c, fsub(comp(v0, c), __ldg(a + c)));
								setcomp(v1, c, fsub(comp(v1, c), __ldg(a + c + Lo.count)));
								setcomp(v2, c, fsub(comp(v2, c), __ldg(a + c + 2ll * Lo.count)));
							}
					}
				}
			}
			if (!halo_row && kk < ke) {
				const long long o = (long long)kk * p.plane + row;
				if (!xsc) { // x-slab chunks are stored by k_xslab_EH
					st4(p.Vd + o, v0);
					st4(p.Vd + p.comp + o, v1);
					st4(p.Vd + 2 * p.comp + o, v2);
				}
			}
			if (hcol) {
				// V1, V2 of cell (xe, j, kk): engine.cpp:148-166 with the x-1 neighbour = my last cell
				const unsigned ex = reinterpret_cast<const IdxT*>(sX)[rV * FT_W + 128];
				const float xi0 = sI[oI + 4], xi1 = sI[cI + oI + 4], xi2 = sI[2 * cI + oI + 4], xi0jm = sI[oIm + 4];
				const float c1x = fadd(fsub(fsub(xi0, hcI0), xi2), i2c.w);
				const float c2x = fadd(fsub(fsub(xi1, i1c.w), xi0), xi0jm);
				const float4 Ax = __ldg(p.eA + ex), Bx = __ldg(p.eB + ex);
				const float b1 = sV[cV + oV + 4], b2 = sV[2 * cV + oV + 4];
				const float n1 = leap(b1, Ax.y, Bx.y, c1x), n2 = leap(b2, Ax.z, Bx.z, c2x);
				nV1 = shx ? b1 : n1;
				nV2 = shx ? b2 : n2;
				if (LOR) {
#pragma unroll
					for (int o = 0; o < LOR_FUSED_MAX; ++o) {
						const LorFusedOrder& Lo = P.lor.o[o];
						const int d = xe - (lorR[o].x & 0xffff), n = lorR[o].x >> 16;
						if (o < P.lor.nord && Lo.ade_v && (unsigned)d < (unsigned)n) {
							const float* a = Lo.ade_v + lorR[o].y + d;
							nV1 = fsub(nV1, __ldg(a + Lo.count));
							nV2 = fsub(nV2, __ldg(a + 2ll * Lo.count));
						}
					}
				}
				nI0 = xi0;
			}
		}
		xV0[kk % 3][ty][lane] = v0;
		xV2[kk % 3][ty][lane] = v2;
		__syncthreads();
		// every thread has taken what it needs of stage s into registers: refill it
		if (producer && kk + STAGES <= e_last) issue(kk + STAGES);

		// ------------------------------------------------------------ H_new(kk-1)
		const int k = kk - 1;
		float r1 = __shfl_down_sync(0xffffffffu, ek1.x, 1);
		float r2 = __shfl_down_sync(0xffffffffu, ek2.x, 1);
		if (lane == 31) { r1 = hcV1; r2 = hcV2; }
		if (k >= kb && !halo_row && active) {
			const long long oh = (long long)k * p.plane + row;
			float4 c0 = hk0, c1 = hk1, c2 = hk2;
			if (k < he && j < p.ny - 1 && !shk && !xsk) { // H of shell / x-slab cells: not updated here
				const float4 v0jp = xV0[k % 3][ty + 1][lane], v2jp = xV2[k % 3][ty + 1][lane];
				const float4 v1xp = make_float4(ek1.y, ek1.z, ek1.w, r1);
				const float4 v2xp = make_float4(ek2.y, ek2.z, ek2.w, r2);
				const bool uni = (ek_idx[0] == ek_idx[1]) & (ek_idx[1] == ek_idx[2]) & (ek_idx[2] == ek_idx[3]);
				float4 A = __ldg(p.hA + ek_idx[0]), B = __ldg(p.hB + ek_idx[0]);
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					if (c > 0 && !uni) { A = __ldg(p.hA + ek_idx[c]); B = __ldg(p.hB + ek_idx[c]); }
					if (ic + c < p.nx - 1) {
						const float curl0 = fadd(fsub(fsub(comp(ek2, c), comp(v2jp, c)), comp(ek1, c)), comp(v1, c));
						const float curl1 = fadd(fsub(fsub(comp(ek0, c), comp(v0, c)), comp(ek2, c)), comp(v2xp, c));
						const float curl2 = fadd(fsub(fsub(comp(ek1, c), comp(v1xp, c)), comp(ek0, c)), comp(v0jp, c));
						setcomp(c0, c, leap(comp(c0, c), A.x, B.x, curl0));
						setcomp(c1, c, leap(comp(c1, c), A.y, B.y, curl1));
						setcomp(c2, c, leap(comp(c2, c), A.z, B.z, curl2));
					}
				}
			}
			if (LOR && k < he) {
				// Apply2Current on the planes this kernel updates (the slab's top plane: list kernel after update_H_top);
				// cells whose H is never updated carry a zero ADE
#pragma unroll
				for (int o = 0; o < LOR_FUSED_MAX; ++o) {
					const LorFusedOrder& Lo = P.lor.o[o];
					const int d = ic - (lorRk[o].x & 0xffff), n = lorRk[o].x >> 16;
					if (o < P.lor.nord && Lo.ade_i && d + 3 >= 0 && d < n) {
						const float* a = Lo.ade_i + lorRk[o].y + d;
#pragma unroll
						for (int c = 0; c < 4; ++c)
							if ((unsigned)(d + c) < (unsigned)n) {
								setcomp(c0, c, fsub(comp(c0, c), __ldg(a + c)));
								setcomp(c1, c, fsub(comp(c1, c), __ldg(a + c + Lo.count)));
								setcomp(c2, c, fsub(comp(c2, c), __ldg(a + c + 2ll * Lo.count)));
							}
					}
				}
			}
			if (k < p.kHc1) {
				if (!xsk) {
					st4(p.Id + oh, c0);
					st4(p.Id + p.comp + oh, c1);
					st4(p.Id + 2 * p.comp + oh, c2);
				}
			}
		}
		// rotate: plane kk becomes "k"
		ek0 = v0; ek1 = v1; ek2 = v2;
		hk0 = i0c; hk1 = i1c; hk2 = i2c;
#pragma unroll
		for (int c = 0; c < 4; ++c) ek_idx[c] = e[c];
		hcV1 = nV1; hcV2 = nV2; hcI0 = nI0;
		shk = sh; xsk = xsc;
		if (LOR)
			for (int o = 0; o < LOR_FUSED_MAX; ++o) lorRk[o] = lorR[o];
	}
	const int k = e_last;
	if (k == ke - 1 && k >= kb && !halo_row && active && k < p.kHc1) {
		const long long oh = (long long)k * p.plane + row;
		if (!xsk) {
			st4(p.Id + oh, hk0);
			st4(p.Id + p.comp + oh, hk1);
			st4(p.Id + 2 * p.comp + oh, hk2);
		}
	}
}
