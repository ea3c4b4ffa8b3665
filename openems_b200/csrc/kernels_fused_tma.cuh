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

// UPML box that is thin in x and lies at the low / high end of the mesh ("x slab"): its cells are
// updated INSIDE the one-pass kernel by the end tiles, one cell per lane (lanes 0-15: low slab,
// 16-31: high slab), from the staged tile -- as a shell launch the same cells cost 64-byte
// pieces of every 4 KB row, which HBM serves at a fraction of its streaming rate.
struct XSlab {
	int n0;              // lines in x, <= 16 (0: no such slab)
	int x_first;         // first line
	int s1, n1, s2, n2;  // y range, local z range
	long long cs;        // flux component stride
	const float* fVs;    // voltage flux of timestep n, component 0 (ping-pong like the fields: the
	float* fVd;          //   halo-row warp and the chunk's extra plane recompute E of cells other blocks own)
	float* fI;           // current flux, in place (every H is computed exactly once)
};

struct alignas(64) FusedTmaParams {
	XSlab xs[2];
	int bx_off, bx_stride; // x tile of a block = bx_off + blockIdx.x * bx_stride
	CUtensorMap mI;   // H_old of the source set: (x, y, z, component) float, box 136 x 9 x 1 x 3
	CUtensorMap mV;   // E_old of the source set (shell cells: already E_new), box 136 x 8 x 1 x 3
	CUtensorMap mX;   // operator index (x, y, z), box 136 x 8 x 1
	const float4 *eP0, *eP1, *eP2, *hP0, *hP1, *hP2; // UPML auxiliary tables (x slabs)
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
	return STAGES * FtStage<IdxT>::BYTES + 3 * 3 * (FUSED_TY + 1) * 32 * 16 + 64; // stages, E_new ring (V0, V2, V1), mbarriers
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

// lane that holds the x-slab result of mesh line x (-1: x is in no x slab)
__device__ __forceinline__ int xslab_lane(const FusedTmaParams& P, int x)
{
	const int a = x - P.xs[0].x_first, b = x - P.xs[1].x_first;
	if ((unsigned)a < (unsigned)P.xs[0].n0) return a;
	if ((unsigned)b < (unsigned)P.xs[1].n0) return 16 + b;
	return -1;
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

template <typename IdxT, bool HAS_PML, int STAGES, bool XS>
__global__ void __launch_bounds__(32 * (FUSED_TY + 1), FT_MIN_BLOCKS) k_fused_tma(const __grid_constant__ FusedTmaParams P)
{
	extern __shared__ __align__(128) unsigned char ft_smem[];
	typedef FtStage<IdxT> ST;
	const FusedParams& p = P.f;
	float4 (*xV0)[FUSED_TY + 1][32] = reinterpret_cast<float4 (*)[FUSED_TY + 1][32]>(ft_smem + STAGES * ST::BYTES);
	float4 (*xV2)[FUSED_TY + 1][32] = xV0 + 3;
	float4 (*xV1)[FUSED_TY + 1][32] = xV0 + 6; // only the x-slab lanes read V1 from the ring
	const uint32_t bar0 = smem_u32(ft_smem + STAGES * ST::BYTES + 3 * 3 * (FUSED_TY + 1) * 32 * 16);

	const int lane = threadIdx.x, ty = threadIdx.y;
	const int x0 = (P.bx_off + (int)blockIdx.x * P.bx_stride) * 128, j0 = blockIdx.y * FUSED_TY;
	const int i0 = x0 + lane * 4;
	const int j = j0 + ty;
	const bool halo_row = ty == FUSED_TY;
	const bool producer = halo_row && lane == 0;
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
	bool shk = false;

	// x-slab lanes: one cell of the row per lane
	const XSlab& X = P.xs[lane >> 4];
	const int xsx = X.x_first + (lane & 15);          // my slab cell
	const int xl = xsx - x0;                          // tile-relative column
	const int ljs = j - X.s1;
	// block-uniform: this tile holds x-slab cells (all tiles run the same kernel: launching the end
	// tiles on their own would read 544-byte pieces of 4 KB rows, which HBM serves poorly)
	const bool xs_tile = XS && ((P.xs[0].n0 > 0 && x0 == 0) || (P.xs[1].n0 > 0 && P.xs[1].x_first >= x0 && P.xs[1].x_first < x0 + 128));
	const bool xs_row = xs_tile && (lane & 15) < X.n0 && xl >= 0 && xl < 128 && row_ok && (unsigned)ljs < (unsigned)X.n1;
	float hs0 = 0.0f, hs1 = 0.0f, hs2 = 0.0f;         // H_old(k) of the slab cell
	unsigned es_idx = 0;
	bool xs_k = false;                                // slab cell updated as UPML at plane k
	if (xs_row) {
		const int km = kb - (kb > 0);
		const long long o = (long long)km * p.plane + (long long)j * p.pitch + xsx;
		hs0 = p.Is[o];
		hs1 = p.Is[p.comp + o];
	}

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

	for (int kk = kb; kk <= e_last; ++kk) {
		// ------------------------------------------------------------ E_new(kk)
		const int s = (kk - kb) % STAGES;
		// x-slab lanes: the flux loads do not depend on the staged tile -- issue them before the wait
		float fv0 = 0.0f, fv1 = 0.0f, fv2 = 0.0f, fi0 = 0.0f, fi1 = 0.0f, fi2 = 0.0f;
		const bool xs_zin = XS && xs_row && (unsigned)(kk - X.s2) < (unsigned)X.n2;
		const bool xs_h = XS && xs_k && !halo_row && kk - 1 < he && j < p.ny - 1 && xsx < p.nx - 1; // xs_k implies kk-1 >= kb
		const long long foV = ((long long)(kk - X.s2) * X.n1 + ljs) * X.n0 + (lane & 15);
		const long long foI = foV - (long long)X.n1 * X.n0;
		if (xs_zin) { fv0 = X.fVs[foV]; fv1 = X.fVs[foV + X.cs]; fv2 = X.fVs[foV + 2 * X.cs]; }
		if (xs_h) { fi0 = X.fI[foI]; fi1 = X.fI[foI + X.cs]; fi2 = X.fI[foI + 2 * X.cs]; }
		mbar_wait(bar0 + 8 * s, ((kk - kb) / STAGES) & 1);
		const unsigned char* stg = ft_smem + s * ST::BYTES;
		const float* sI = reinterpret_cast<const float*>(stg);
		const float* sV = reinterpret_cast<const float*>(stg + ST::V_OFF);
		const unsigned char* sX = stg + ST::X_OFF;
		const bool sh = HAS_PML && (shb >> (kk - kb) & 1ull), shx = HAS_PML && (shxb >> (kk - kb) & 1ull);
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
		}
		// ---- x-slab cells of this row: UPML update, one cell per lane, merged into the row's float4s
		float h0n = 0.0f, h1n = 0.0f, h2n = 0.0f;
		unsigned eidx = 0;
		bool xs_kk = false;
		if (xs_tile) {
			float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;
			if (xs_row) {
				const int col = 4 + xl;
				h0n = sI[rI * FT_W + col]; h1n = sI[cI + rI * FT_W + col]; h2n = sI[2 * cI + rI * FT_W + col];
				eidx = reinterpret_cast<const IdxT*>(sX)[rV * FT_W + xl];
				if (xs_zin) {
					const float4 A = __ldg(p.eA + eidx);
					if (A.w != 0.0f) {
						const float4 B = __ldg(p.eB + eidx), P0 = __ldg(P.eP0 + eidx), P1 = __ldg(P.eP1 + eidx), P2 = __ldg(P.eP2 + eidx);
						const float j0m = sI[rIm * FT_W + col], j2m = sI[2 * cI + rIm * FT_W + col];
						const float x1m = xsx > 0 ? sI[cI + rI * FT_W + col - 1] : h1n, x2m = xsx > 0 ? sI[2 * cI + rI * FT_W + col - 1] : h2n;
						const float curl0 = fadd(fsub(fsub(h2n, j2m), h1n), hs1);
						const float curl1 = fadd(fsub(fsub(h0n, hs0), h2n), x2m);
						const float curl2 = fadd(fsub(fsub(h1n, x1m), h0n), j0m);
						float f0, f1, f2;
						r0 = leap_pml_oop(sV[rV * FT_W + col], A.x, B.x, curl0, P0.x, P1.x, P2.x, fv0, f0);
						r1 = leap_pml_oop(sV[cV + rV * FT_W + col], A.y, B.y, curl1, P0.y, P1.y, P2.y, fv1, f1);
						r2 = leap_pml_oop(sV[2 * cV + rV * FT_W + col], A.z, B.z, curl2, P0.z, P1.z, P2.z, fv2, f2);
						if (!halo_row && kk < ke) { X.fVd[foV] = f0; X.fVd[foV + X.cs] = f1; X.fVd[foV + 2 * X.cs] = f2; }
						xs_kk = true;
					}
				}
			}
			const unsigned bal = __ballot_sync(0xffffffffu, xs_kk);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const int src = xslab_lane(P, ic + c);
				const int sl = src < 0 ? lane : src;
				const float t0 = __shfl_sync(0xffffffffu, r0, sl), t1 = __shfl_sync(0xffffffffu, r1, sl), t2 = __shfl_sync(0xffffffffu, r2, sl);
				if (src >= 0 && (bal >> src & 1u)) { setcomp(v0, c, t0); setcomp(v1, c, t1); setcomp(v2, c, t2); }
			}
		}
		if (active) {
			if (!halo_row && kk < ke) {
				const long long o = (long long)kk * p.plane + row;
				st4(p.Vd + o, v0);
				st4(p.Vd + p.comp + o, v1);
				st4(p.Vd + 2 * p.comp + o, v2);
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
				nI0 = xi0;
			}
		}
		xV0[kk % 3][ty][lane] = v0;
		xV2[kk % 3][ty][lane] = v2;
		if (xs_tile) xV1[kk % 3][ty][lane] = v1;
		__syncthreads();
		// every thread has taken what it needs of stage s into registers: refill it
		if (producer && kk + STAGES <= e_last) issue(kk + STAGES);

		// ------------------------------------------------------------ H_new(kk-1)
		const int k = kk - 1;
		float r1 = __shfl_down_sync(0xffffffffu, ek1.x, 1);
		float r2 = __shfl_down_sync(0xffffffffu, ek2.x, 1);
		if (lane == 31) { r1 = hcV1; r2 = hcV2; }
		// x-slab cells: H_new(k) from the E_new ring
		float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f;
		unsigned balh = 0;
		if (xs_tile && !halo_row && k >= kb) { // warp-uniform
			bool hb = false;
			if (xs_h) {
				const float* a0 = reinterpret_cast<const float*>(&xV0[k % 3][ty][0]) + xl;
				const float* a1 = reinterpret_cast<const float*>(&xV1[k % 3][ty][0]) + xl;
				const float* a2 = reinterpret_cast<const float*>(&xV2[k % 3][ty][0]) + xl;
				const float w0 = a0[0], w1 = a1[0], w2 = a2[0];
				const float w0jp = a0[32 * 4], w2jp = a2[32 * 4];   // row ty+1 of the ring
				const float w1xp = a1[1], w2xp = a2[1];
				const float w0n = reinterpret_cast<const float*>(&xV0[kk % 3][ty][0])[xl], w1n = reinterpret_cast<const float*>(&xV1[kk % 3][ty][0])[xl];
				const float curl0 = fadd(fsub(fsub(w2, w2jp), w1), w1n);
				const float curl1 = fadd(fsub(fsub(w0, w0n), w2), w2xp);
				const float curl2 = fadd(fsub(fsub(w1, w1xp), w0), w0jp);
				const float4 A = __ldg(p.hA + es_idx), B = __ldg(p.hB + es_idx), P0 = __ldg(P.hP0 + es_idx), P1 = __ldg(P.hP1 + es_idx), P2 = __ldg(P.hP2 + es_idx);
				float f0, f1, f2;
				q0 = leap_pml_oop(hs0, A.x, B.x, curl0, P0.x, P1.x, P2.x, fi0, f0);
				q1 = leap_pml_oop(hs1, A.y, B.y, curl1, P0.y, P1.y, P2.y, fi1, f1);
				q2 = leap_pml_oop(hs2, A.z, B.z, curl2, P0.z, P1.z, P2.z, fi2, f2);
				X.fI[foI] = f0; X.fI[foI + X.cs] = f1; X.fI[foI + 2 * X.cs] = f2;
				hb = true;
			}
			balh = __ballot_sync(0xffffffffu, hb);
		}
		float4 m0 = make_float4(0, 0, 0, 0), m1 = m0, m2 = m0;
		unsigned mm = 0;
		if (xs_tile && !halo_row && k >= kb) {
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const int src = xslab_lane(P, ic + c);
				const int sl = src < 0 ? lane : src;
				setcomp(m0, c, __shfl_sync(0xffffffffu, q0, sl));
				setcomp(m1, c, __shfl_sync(0xffffffffu, q1, sl));
				setcomp(m2, c, __shfl_sync(0xffffffffu, q2, sl));
				if (src >= 0 && (balh >> src & 1u)) mm |= 1u << c;
			}
		}
		if (k >= kb && !halo_row && active) {
			const long long oh = (long long)k * p.plane + row;
			float4 c0 = hk0, c1 = hk1, c2 = hk2;
			if (k < he && j < p.ny - 1 && !shk) { // H of shell cells: copied through here, updated by k_shell_H
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
			if (XS && mm) {
#pragma unroll
				for (int c = 0; c < 4; ++c)
					if (mm >> c & 1u) { setcomp(c0, c, comp(m0, c)); setcomp(c1, c, comp(m1, c)); setcomp(c2, c, comp(m2, c)); }
			}
			if (k < p.kHc1) {
				st4(p.Id + oh, c0);
				st4(p.Id + p.comp + oh, c1);
				st4(p.Id + 2 * p.comp + oh, c2);
			}
		}
		if (xs_tile) { hs0 = h0n; hs1 = h1n; hs2 = h2n; es_idx = eidx; xs_k = xs_kk; }
		// rotate: plane kk becomes "k"
		ek0 = v0; ek1 = v1; ek2 = v2;
		hk0 = i0c; hk1 = i1c; hk2 = i2c;
#pragma unroll
		for (int c = 0; c < 4; ++c) ek_idx[c] = e[c];
		hcV1 = nV1; hcV2 = nV2; hcI0 = nI0;
		shk = sh;
	}
	const int k = e_last;
	if (k == ke - 1 && k >= kb && !halo_row && active && k < p.kHc1) {
		const long long oh = (long long)k * p.plane + row;
		st4(p.Id + oh, hk0);
		st4(p.Id + p.comp + oh, hk1);
		st4(p.Id + 2 * p.comp + oh, hk2);
	}
}
