// kernels_xslab_tma.cuh -- one-pass timestep of the "x slabs" (UPML boxes that are thin in x and sit at the low /
// high end of the mesh), TMA-staged.
//
// Why these boxes get their own kernel: a 9-line slab uses 36..48 bytes of every 4 KB row; whatever reads them,
// HBM delivers whole lines, so as two shell launches (k_shell_E / k_shell_H) they cost 1.5 ms of the 11.1 ms step
// for 1.8 % of the cells (profiles/experiments_r01.md #14).  k_xslab_EH (kernels_xslab.cuh) already reads the lines
// once per step, but with scalar loads it keeps too few bytes in flight (3.0 TB/s, #15).  This version
//   * owns a window of 16 lines = 64 bytes = two whole 32-byte sectors per row (lines that lie in no UPML box get
//     the plain leapfrog), so every store fills whole sectors and the big kernel needs no store masks: it treats the
//     window like any other cells (plain leapfrog, wrong for the UPML cells) and this kernel, launched after it,
//     overwrites the window in the destination set.  Nothing outside the window reads a window value of the
//     destination set before that (H(i) needs E(i+1): only window cells read window cells; the first line right
//     of a high window is a plain cell, computed identically by both kernels);
//   * stages its inputs like k_fused_tma: three bulk tensor copies per plane (H 24 x 17 x 3, E 24 x 16 x 3, index
//     24 x 16) into a ring of STAGES shared-memory stages with one mbarrier each, so several planes of every block
//     are in flight and HBM latency is covered by the ring, not by occupancy;
//   * one float4 chunk per thread, 4 lanes side by side in x, 16 rows (15 + halo row) = 64 threads per block.  Other
//     thread mappings were tried and are slower (profiles/experiments_r02.md #9): a warp per chunk column inside one
//     block (straight-line UPML / plain code, but the plain warp waits at the per-plane barrier: 1.11 ms), a warp per
//     chunk column as its own block (no barrier, but the columns drift apart in z and the 32-byte sectors they share
//     are fetched and written half-filled several times: 2.9 ms), one interleaved coefficient table (0.94 ms);
//     same E_new(kk) -> ring -> H_new(kk-1) schedule, helpers and roundings as everywhere else: bit-identical.
// The voltage flux is ping-ponged with the field sets (E of halo rows / the extra plane is recomputed by
// neighbouring blocks), the current flux is updated in place.  Cells of the window that belong to ANOTHER box
// (rows / planes where a y or z box takes over) or lie in its chunk-aligned footprint were updated in place by that
// box's k_shell_E before: their E and H are passed on unchanged (that box's k_shell_H follows).
#pragma once
#include "kernels_fused_tma.cuh"
#include "kernels_xslab.cuh"

#define XT_WL 4                        // float4 chunks per row: the 16-line window
#define XT_TY 15                       // rows a block owns (+1 halo row)
#define XT_ROWS (XT_TY + 1)
#define XT_THREADS (XT_WL * XT_ROWS)   // 64
#define XT_W (XT_WL * 4 + 8)           // staged row: ws-4 .. ws+19
#define XT_ROWS_I (XT_TY + 2)          // H rows j0-1 .. j0+TY
#define XT_ROWS_V (XT_TY + 1)          // E / index rows j0 .. j0+TY
#ifndef XT_STAGES
#define XT_STAGES 2      // ring depth / resident blocks: 2 / 6 measured best at 1024^3 (0.97 ms; 4 / 4: 1.08, 3 / 6: 1.02, 2 / 8: 1.05 -- spills)
#endif
#ifndef XT_MIN_BLOCKS
#define XT_MIN_BLOCKS 6
#endif

struct alignas(64) XTmaParams {
	CUtensorMap mI;   // H_old of the source set, box 24 x 17 x 1 x 3
	CUtensorMap mV;   // E_old of the source set, box 24 x 16 x 1 x 3
	CUtensorMap mX;   // operator index, box 24 x 16 x 1
	XSlabParams x;
};

template <typename IdxT> struct XtStage {
	static constexpr int I_BYTES = 3 * XT_ROWS_I * XT_W * 4;
	static constexpr int V_BYTES = 3 * XT_ROWS_V * XT_W * 4;
	static constexpr int X_BYTES = XT_ROWS_V * XT_W * (int)sizeof(IdxT);
	static constexpr int V_OFF = (I_BYTES + 127) / 128 * 128;
	static constexpr int X_OFF = V_OFF + (V_BYTES + 127) / 128 * 128;
	static constexpr int BYTES = X_OFF + (X_BYTES + 127) / 128 * 128;
	static constexpr int TX = I_BYTES + V_BYTES + X_BYTES;
};
template <typename IdxT, int STAGES> constexpr int xt_smem_bytes()
{
	return STAGES * XtStage<IdxT>::BYTES + 2 * 3 * XT_ROWS * XT_WL * 16 + 64;
}

template <typename IdxT, int STAGES>
__global__ void __launch_bounds__(XT_THREADS, XT_MIN_BLOCKS) k_xslab_tma(const __grid_constant__ XTmaParams P)
{
	PDL_PROLOGUE();
	extern __shared__ __align__(128) unsigned char xt_smem[];
	typedef XtStage<IdxT> ST;
	const XSlabParams& p = P.x;
	const XSlabBox& B = p.box[blockIdx.z];
	float4 (*xV0)[XT_ROWS][XT_WL] = reinterpret_cast<float4 (*)[XT_ROWS][XT_WL]>(xt_smem + STAGES * ST::BYTES);
	float4 (*xV2)[XT_ROWS][XT_WL] = xV0 + 3;
	const uint32_t bar0 = smem_u32(xt_smem + STAGES * ST::BYTES + 2 * 3 * XT_ROWS * XT_WL * 16);

	const int tid = threadIdx.x;
	const int lx = tid & (XT_WL - 1), ty = tid / XT_WL;
	const int ws = B.w0, j0 = p.jb + blockIdx.x * XT_TY;
	const int ic = ws + lx * 4;
	const int j = j0 + ty;
	const bool halo_row = ty == XT_TY || j >= p.je;
	const bool producer = tid == XT_THREADS - 1;
	const int kb = p.kE0 + blockIdx.y * p.zchunk;
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
		const uint32_t dst = smem_u32(xt_smem + s * ST::BYTES);
		mbar_expect_tx(bar, ST::TX);
		tma_load_4d(dst, &P.mI, ws - 4, j0 - 1, kk, 0, bar);
		tma_load_4d(dst + ST::V_OFF, &P.mV, ws - 4, j0, kk, 0, bar);
		tma_load_3d(dst + ST::X_OFF, &P.mX, ws, j0, kk, bar);
	};
	if (producer)
		for (int s = 0; s < STAGES && kb + s <= e_last; ++s) issue(kb + s);

	const bool row_ok = j < p.ny;
	const bool active = row_ok;                      // the window lies inside the pitch (host check)
	const int jc = row_ok ? j : p.ny - 1;
	const long long row = (long long)jc * p.pitch + ic;
	const int rI = ty + 1, rIm = (j > 0) ? ty : ty + 1, rV = ty;
	const int xe = ic + 4;
	const bool hcol = lx == XT_WL - 1 && !halo_row && active && xe < p.nx;
	const bool row_in = jc >= B.oj0 && jc < B.oj1;
	const long long fplane = (long long)B.n1 * B.bn0;
	const long long frow = (long long)(jc - B.s1) * B.bn0 + (ic - B.bs0);   // + (k - s2) * fplane + c
	// flux of this chunk at plane kk (voltages: source / destination set) and kk-1 (currents), component 0;
	// components cs apart
	const long long fo0 = (long long)(kb - B.s2) * fplane + frow;
	const float* fvs = B.fVs + fo0;
	float* fvd = B.fVd + fo0;
	float* fic = B.fI + fo0 - fplane;
	const long long cs = B.cs;
	// lines of this chunk that lie in the box (bit c)
	unsigned inbox = 0;
#pragma unroll
	for (int c = 0; c < 4; ++c)
		if ((unsigned)(ic + c - B.bs0) < (unsigned)B.bn0 && ic + c < p.nx) inbox |= 1u << c;
	if (!row_in || !row_ok) inbox = 0;
	// planes of the march (bit kk-kb) where this chunk / the chunk right of the window lies in the chunk-aligned
	// footprint of a box the shell launches update (the march has at most 63 planes + the extra one: host)
	unsigned long long shb = 0, shxb = 0;
	for (int b = 0; b < p.nsh; ++b) {
		if ((unsigned)(jc - p.sh[b].j0) >= (unsigned)p.sh[b].jn) continue;
		const bool mine = (unsigned)((ic >> 2) - p.sh[b].c0) < (unsigned)p.sh[b].cn;
		const bool right = (unsigned)((xe >> 2) - p.sh[b].c0) < (unsigned)p.sh[b].cn;
		if (!mine && !right) continue;
		const int a = max(p.sh[b].k0, kb) - kb, z = min(p.sh[b].k0 + p.sh[b].kn, e_last + 1) - kb;
		if (z <= a) continue;
		const unsigned long long m = (z - a >= 64 ? ~0ull : ((1ull << (z - a)) - 1ull)) << a;
		if (mine) shb |= m;
		if (right) shxb |= m;
	}

	float4 ek0 = make_float4(0, 0, 0, 0), ek1 = ek0, ek2 = ek0;
	float4 hk0 = ek0, hk1 = ek0, hk2 = ek0;
	unsigned ek_idx[4] = {0, 0, 0, 0};
	float hcV1 = 0.0f, hcV2 = 0.0f, hcI0 = 0.0f;
	unsigned pml_k = 0;       // cells of plane k (previous iteration) that are UPML cells of this box
	unsigned frn_k = 0;       // cells of plane k that another box's shell launches update
	if (active) {
		const int km = kb - (kb > 0);
		const long long o = (long long)km * p.plane + row;
		hk0 = ld4(p.Is + o);
		hk1 = ld4(p.Is + p.comp + o);
		if (hcol) hcI0 = p.Is[o + 4];
	}

	for (int kk = kb; kk <= e_last; ++kk) {
		const bool plane_in = kk >= B.ok0 && kk < B.ok1;
		// ---- flux of timestep n of this chunk's box cells: plane kk (voltage flux) and plane kk-1 (current flux),
		// issued before the wait for the staged plane
		float fv[3][4], fi[3][4];
		const unsigned box_kk = plane_in ? inbox : 0u;
#pragma unroll
		for (int c = 0; c < 4; ++c) {
#pragma unroll
			for (int n = 0; n < 3; ++n) {
				fv[n][c] = (box_kk >> c & 1u) ? __ldg(fvs + c + n * cs) : 0.0f;
				fi[n][c] = (pml_k >> c & 1u) ? fic[c + n * cs] : 0.0f;
			}
		}
		// ------------------------------------------------------------ E_new(kk)
		const int s = (kk - kb) % STAGES;
		mbar_wait(bar0 + 8 * s, ((kk - kb) / STAGES) & 1);
		const unsigned char* stg = xt_smem + s * ST::BYTES;
		const float* sI = reinterpret_cast<const float*>(stg);
		const float* sV = reinterpret_cast<const float*>(stg + ST::V_OFF);
		const unsigned char* sX = stg + ST::X_OFF;
		const int cI = XT_ROWS_I * XT_W, cV = XT_ROWS_V * XT_W;
		const int oI = rI * XT_W + 4 + lx * 4, oIm = rIm * XT_W + 4 + lx * 4, oV = rV * XT_W + 4 + lx * 4;
		unsigned e[4];
		SIdx4<IdxT>::load(sX + rV * XT_W * (int)sizeof(IdxT), lx, e);
		const float4 i0c = ld4(sI + oI), i1c = ld4(sI + cI + oI), i2c = ld4(sI + 2 * cI + oI);
		const float4 i0jm = ld4(sI + oIm), i2jm = ld4(sI + 2 * cI + oIm);
		float4 v0 = ld4(sV + oV), v1 = ld4(sV + cV + oV), v2 = ld4(sV + 2 * cV + oV);
		float l1 = __shfl_up_sync(0xffffffffu, i1c.w, 1, XT_WL);
		float l2 = __shfl_up_sync(0xffffffffu, i2c.w, 1, XT_WL);
		if (lx == 0) {
			if (ic > 0) { l1 = sI[cI + oI - 1]; l2 = sI[2 * cI + oI - 1]; }
			else { l1 = i1c.x; l2 = i2c.x; }
		}
		const float4 i1xm = make_float4(l1, i1c.x, i1c.y, i1c.z);
		const float4 i2xm = make_float4(l2, i2c.x, i2c.y, i2c.z);
		float nV1 = 0.0f, nV2 = 0.0f, nI0 = 0.0f;
		unsigned pml_kk = 0;
		const bool inshell = shb >> (kk - kb) & 1ull;
		unsigned frn_kk = 0;
		if (active) {
			float fn[3][4];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const float4 A = __ldg(p.eA + e[c]), Bc = __ldg(p.eB + e[c]);
				const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(hk1, c));
				const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(hk0, c)), comp(i2c, c)), comp(i2xm, c));
				const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
				const bool flag = A.w != 0.0f;
				const bool own = flag && (box_kk >> c & 1u);
				// not a cell of this box, but flagged (a cell of another box) or inside another box's chunk-aligned
				// footprint: that box's k_shell_E has updated it in place, its k_shell_H follows
				const bool frn = !own && (inshell || flag);
				if (frn) frn_kk |= 1u << c;
				fn[0][c] = fn[1][c] = fn[2][c] = 0.0f;
				if (own) {
					const float4 P0 = __ldg(p.eP0 + e[c]), P1 = __ldg(p.eP1 + e[c]), P2 = __ldg(p.eP2 + e[c]);
					setcomp(v0, c, leap_pml_oop(comp(v0, c), A.x, Bc.x, curl0, P0.x, P1.x, P2.x, fv[0][c], fn[0][c]));
					setcomp(v1, c, leap_pml_oop(comp(v1, c), A.y, Bc.y, curl1, P0.y, P1.y, P2.y, fv[1][c], fn[1][c]));
					setcomp(v2, c, leap_pml_oop(comp(v2, c), A.z, Bc.z, curl2, P0.z, P1.z, P2.z, fv[2][c], fn[2][c]));
					pml_kk |= 1u << c;
				} else if (!frn) {
					setcomp(v0, c, leap(comp(v0, c), A.x, Bc.x, curl0));
					setcomp(v1, c, leap(comp(v1, c), A.y, Bc.y, curl1));
					setcomp(v2, c, leap(comp(v2, c), A.z, Bc.z, curl2));
				}
			}
			if (!halo_row && kk < ke) {
				const long long o = (long long)kk * p.plane + row;
				st4(p.Vd + o, v0);
				st4(p.Vd + p.comp + o, v1);
				st4(p.Vd + 2 * p.comp + o, v2);
#pragma unroll
				for (int c = 0; c < 4; ++c)
					if (pml_kk >> c & 1u) {
						fvd[c] = fn[0][c]; fvd[c + cs] = fn[1][c]; fvd[c + 2 * cs] = fn[2][c];
					}
			}
			if (hcol) {
				// V1, V2 of cell (xe, j, kk), the first line right of the window: a plain cell, or a cell another box's
				// shell has already updated in place (then its value is taken as it is)
				const unsigned ex = reinterpret_cast<const IdxT*>(sX)[rV * XT_W + XT_WL * 4];
				const float xi0 = sI[oI + 4], xi1 = sI[cI + oI + 4], xi2 = sI[2 * cI + oI + 4], xi0jm = sI[oIm + 4];
				const float c1x = fadd(fsub(fsub(xi0, hcI0), xi2), i2c.w);
				const float c2x = fadd(fsub(fsub(xi1, i1c.w), xi0), xi0jm);
				const float4 Ax = __ldg(p.eA + ex), Bx = __ldg(p.eB + ex);
				const float b1 = sV[cV + oV + 4], b2 = sV[2 * cV + oV + 4];
				const bool shx = Ax.w != 0.0f || (shxb >> (kk - kb) & 1ull);
				nV1 = shx ? b1 : leap(b1, Ax.y, Bx.y, c1x);
				nV2 = shx ? b2 : leap(b2, Ax.z, Bx.z, c2x);
				nI0 = xi0;
			}
		}
		xV0[kk % 3][ty][lx] = v0;
		xV2[kk % 3][ty][lx] = v2;
		__syncthreads();
		if (producer && kk + STAGES <= e_last) issue(kk + STAGES);

		// ------------------------------------------------------------ H_new(kk-1)
		const int k = kk - 1;
		float r1 = __shfl_down_sync(0xffffffffu, ek1.x, 1, XT_WL);
		float r2 = __shfl_down_sync(0xffffffffu, ek2.x, 1, XT_WL);
		if (lx == XT_WL - 1) { r1 = hcV1; r2 = hcV2; }
		if (k >= kb && !halo_row && active) {
			const long long oh = (long long)k * p.plane + row;
			float4 c0 = hk0, c1 = hk1, c2 = hk2;
			if (k < he && j < p.ny - 1) {
				const float4 v0jp = xV0[k % 3][ty + 1][lx], v2jp = xV2[k % 3][ty + 1][lx];
				const float4 v1xp = make_float4(ek1.y, ek1.z, ek1.w, r1);
				const float4 v2xp = make_float4(ek2.y, ek2.z, ek2.w, r2);
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					if (ic + c < p.nx - 1 && !(frn_k >> c & 1u)) { // cells of other boxes: H passed on, their k_shell_H follows
						const float4 A = __ldg(p.hA + ek_idx[c]), Bh = __ldg(p.hB + ek_idx[c]);
						const float curl0 = fadd(fsub(fsub(comp(ek2, c), comp(v2jp, c)), comp(ek1, c)), comp(v1, c));
						const float curl1 = fadd(fsub(fsub(comp(ek0, c), comp(v0, c)), comp(ek2, c)), comp(v2xp, c));
						const float curl2 = fadd(fsub(fsub(comp(ek1, c), comp(v1xp, c)), comp(ek0, c)), comp(v0jp, c));
						if (pml_k >> c & 1u) {
							const float4 P0 = __ldg(p.hP0 + ek_idx[c]), P1 = __ldg(p.hP1 + ek_idx[c]), P2 = __ldg(p.hP2 + ek_idx[c]);
							float f0, f1, f2;
							setcomp(c0, c, leap_pml_oop(comp(c0, c), A.x, Bh.x, curl0, P0.x, P1.x, P2.x, fi[0][c], f0));
							setcomp(c1, c, leap_pml_oop(comp(c1, c), A.y, Bh.y, curl1, P0.y, P1.y, P2.y, fi[1][c], f1));
							setcomp(c2, c, leap_pml_oop(comp(c2, c), A.z, Bh.z, curl2, P0.z, P1.z, P2.z, fi[2][c], f2));
							float* f = fic + c;
							f[0] = f0; f[cs] = f1; f[2 * cs] = f2;
						} else {
							setcomp(c0, c, leap(comp(c0, c), A.x, Bh.x, curl0));
							setcomp(c1, c, leap(comp(c1, c), A.y, Bh.y, curl1));
							setcomp(c2, c, leap(comp(c2, c), A.z, Bh.z, curl2));
						}
					}
				}
			}
			if (k < p.kHc1) {
				st4(p.Id + oh, c0);
				st4(p.Id + p.comp + oh, c1);
				st4(p.Id + 2 * p.comp + oh, c2);
			}
		}
		ek0 = v0; ek1 = v1; ek2 = v2;
		hk0 = i0c; hk1 = i1c; hk2 = i2c;
#pragma unroll
		for (int c = 0; c < 4; ++c) ek_idx[c] = e[c];
		hcV1 = nV1; hcV2 = nV2; hcI0 = nI0;
		pml_k = pml_kk; frn_k = frn_kk;
		fvs += fplane; fvd += fplane; fic += fplane;
	}
	const int k = e_last;
	if (k == ke - 1 && k >= kb && !halo_row && active && k < p.kHc1) {
		const long long oh = (long long)k * p.plane + row;
		st4(p.Id + oh, hk0);
		st4(p.Id + p.comp + oh, hk1);
		st4(p.Id + 2 * p.comp + oh, hk2);
	}
}
