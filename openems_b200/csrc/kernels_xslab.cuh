// kernels_xslab.cuh -- one-pass timestep of the "x slabs": UPML boxes that are thin in x and sit at the
// low / high end of the mesh.
//
// As shell launches (k_shell_E / k_shell_H) these boxes are the expensive part of the UPML shell: a 9-line
// slab uses 48 bytes of every 4 KB row, HBM delivers a 128-byte line for each, and the two half-steps fetch
// the same lines twice (profiles/experiments_r01.md #14).  k_xslab_EH reads them once: E and H of the
// slab's float4 chunks in one pass, out of place (source set -> destination set) like k_fused_tma, which
// leaves these chunks alone (its store masks, kernels_fused_tma.cuh).  The voltage flux is ping-ponged with
// the field sets because E of halo rows / the extra plane is recomputed by neighbouring blocks; the current
// flux is updated in place (every H is computed once).
//
// Block = 16 lines in x (the slab's chunks + the halo column on the inner side) x 16 rows (15 + halo row),
// one cell per thread, marching a z chunk upwards with the same schedule as the big kernel:
//   E_new(kk) -> 3-slot shared-memory ring -> __syncthreads -> H_new(kk-1).
// Cells of the slab's chunks that lie in no UPML box get the plain leapfrog; cells of another box (rows /
// planes where a y or z box takes over) were updated in place by that box's k_shell_E before this kernel
// runs: their E and H are passed on to the destination set unchanged (that box's k_shell_H follows).
// Same helpers and roundings as everywhere else: bit-identical to the two-pass schedule.
#pragma once
#include "kernels_fused.cuh"

#define XSLAB_ROWS 15 // rows a block owns (+1 halo row = 16 threads in y)
#ifndef XSLAB_PF
#define XSLAB_PF 3     // planes ahead that are pulled into L2
#endif

struct XSlabBox {
	int w0;              // first line of the 16-line window
	int own0, own1;      // lines stored by this kernel: the box's float4 chunks [own0, own1)
	int bs0, bn0;        // the box: first line, lines in x
	int s1, n1, s2, n2;  // rows / local planes of the box
	int oj0, oj1, ok0, ok1; // k_xslab_tma: rows / local planes of the box this kernel updates (the rest: shell launches)
	long long cs;        // flux component stride
	const float* fVs;    // voltage flux of timestep n   (component 0)
	float* fVd;          // voltage flux of timestep n+1
	float* fI;           // current flux, in place
};
struct XSlabParams {
	const float* Vs; const float* Is;
	float* Vd; float* Id;
	const void* idx;
	const float4 *eA, *eB, *eP0, *eP1, *eP2;
	const float4 *hA, *hB, *hP0, *hP1, *hP2;
	int nx, ny, nz;      // nz = local planes held
	int pitch;
	long long plane, comp;
	int kE0, kE1, kH1, kHc1; // as in FusedParams
	int jb, je;          // k_xslab_tma: rows stored, as in FusedParams
	int zchunk;
	XSlabBox box[2];
	// chunk-aligned footprints of the boxes the shell launches handle: a cell of the slab's chunks that lies
	// in no box but inside such a footprint was already updated in place by that shell
	int nsh;
	struct { int c0, cn, j0, jn, k0, kn; } sh[OEMS_MAX_PML_BOXES];
};

__device__ __forceinline__ bool xslab_in_shell(const XSlabParams& p, int chunk, int j, int k)
{
	for (int b = 0; b < p.nsh; ++b)
		if ((unsigned)(chunk - p.sh[b].c0) < (unsigned)p.sh[b].cn && (unsigned)(j - p.sh[b].j0) < (unsigned)p.sh[b].jn && (unsigned)(k - p.sh[b].k0) < (unsigned)p.sh[b].kn) return true;
	return false;
}

template <typename IdxT>
__global__ void __launch_bounds__(256) k_xslab_EH(const __grid_constant__ XSlabParams p)
{
	PDL_PROLOGUE();
	__shared__ float rV[3][3][XSLAB_ROWS + 1][16]; // [slot][component][row][x]

	const XSlabBox& B = p.box[blockIdx.z];
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int x = B.w0 + tx;
	const int j = B.s1 + blockIdx.x * XSLAB_ROWS + ty;
	const bool halo_row = ty == XSLAB_ROWS;
	const int kb = max(p.kE0, B.s2) + blockIdx.y * p.zchunk;
	const int kend = min(p.kE1, B.s2 + B.n2);
	const int ke = min(kb + p.zchunk, kend);
	if (kb >= ke) return; // block-uniform
	const int he = min(ke, p.kH1);
	const int e_last = (he == ke && ke < p.nz) ? ke : ke - 1;

	const bool cell_ok = x < p.nx && j < p.ny;                       // a mesh cell
	const bool own_x = x >= B.own0 && x < B.own1;                    // stored by this kernel (if the row / plane is the box's)
	const bool row_in = (unsigned)(j - B.s1) < (unsigned)B.n1;
	const bool box_x = (unsigned)(x - B.bs0) < (unsigned)B.bn0;
	const int xc = min(x, p.nx - 1), jc = min(j, p.ny - 1);
	const int jm = jc - (jc > 0);
	const long long row = (long long)jc * p.pitch + xc;
	const long long rowm = (long long)jm * p.pitch + xc;
	const long long frow = ((long long)(j - B.s1)) * B.bn0 + (x - B.bs0); // + (k - s2) * n1 * bn0

	float hk0, hk1, hk2 = 0.0f;     // H_old(k) of this cell (k-1 for the E update, k for the H update)
	unsigned ek_idx = 0;
	bool ownpml_k = false, foreign_k = false;
	{
		const int km = kb - (kb > 0);
		const long long o = (long long)km * p.plane + row;
		hk0 = p.Is[o];
		hk1 = p.Is[p.comp + o];
	}
	// software pipeline: the loads of plane kk+1 are issued before plane kk is computed (the block's loop is
	// latency bound otherwise: a 16 x 16 block has little else to hide HBM latency with)
	unsigned n_e; float n_h0, n_h1, n_h2, n_j0m, n_j2m, n_v0, n_v1, n_v2, n_l1 = 0.0f, n_l2 = 0.0f;
	{
		const long long o = (long long)kb * p.plane + row, om = (long long)kb * p.plane + rowm;
		n_e = reinterpret_cast<const IdxT*>(p.idx)[o];
		n_h0 = p.Is[o]; n_h1 = p.Is[p.comp + o]; n_h2 = p.Is[2 * p.comp + o];
		n_j0m = p.Is[om]; n_j2m = p.Is[2 * p.comp + om];
		n_v0 = p.Vs[o]; n_v1 = p.Vs[p.comp + o]; n_v2 = p.Vs[2 * p.comp + o];
		if (tx == 0 && xc > 0) { n_l1 = p.Is[p.comp + o - 1]; n_l2 = p.Is[2 * p.comp + o - 1]; }
	}
	const bool geo = box_x && row_in && cell_ok; // UPML cell of this box if the plane is the box's and the tuple is flagged
	const long long fplane = (long long)B.n1 * B.bn0;

	for (int kk = kb; kk <= e_last; ++kk) {
		const long long o = (long long)kk * p.plane + row;
		// pull the lines of plane kk+XSLAB_PF into L2 (one lane per row: a row of the window is one 64-byte piece,
		// HBM delivers its 128-byte line): the block has too few loads in flight to cover HBM latency otherwise
		if (tx == 0 && kk + XSLAB_PF <= e_last) {
			const long long of = o + (long long)XSLAB_PF * p.plane;
			prefetch_l2(p.Is + of); prefetch_l2(p.Is + p.comp + of); prefetch_l2(p.Is + 2 * p.comp + of);
			prefetch_l2(p.Vs + of); prefetch_l2(p.Vs + p.comp + of); prefetch_l2(p.Vs + 2 * p.comp + of);
			prefetch_l2(reinterpret_cast<const IdxT*>(p.idx) + of);
			if (row_in && (unsigned)(kk + XSLAB_PF - B.s2) < (unsigned)B.n2) {
				const long long ff = (long long)(kk + XSLAB_PF - B.s2) * fplane + ((long long)(j - B.s1)) * B.bn0;
				prefetch_l2(B.fVs + ff); prefetch_l2(B.fVs + ff + B.cs); prefetch_l2(B.fVs + ff + 2 * B.cs);
				prefetch_l2(B.fI + ff); prefetch_l2(B.fI + ff + B.cs); prefetch_l2(B.fI + ff + 2 * B.cs);
			}
		}
		const unsigned e = n_e;
		const float h0 = n_h0, h1 = n_h1, h2 = n_h2, j0m = n_j0m, j2m = n_j2m;
		float v0 = n_v0, v1 = n_v1, v2 = n_v2;
		float l1 = n_l1, l2 = n_l2;
		// ---- everything this iteration needs from memory, issued up front
		const bool plane_in = (unsigned)(kk - B.s2) < (unsigned)B.n2;
		const bool plane_k_in = (unsigned)(kk - 1 - B.s2) < (unsigned)B.n2;
		const long long fo = (long long)(kk - B.s2) * fplane + frow;
		float fv0 = 0.0f, fv1 = 0.0f, fv2 = 0.0f, fi0 = 0.0f, fi1 = 0.0f, fi2 = 0.0f;
		if (geo && plane_in) { fv0 = B.fVs[fo]; fv1 = B.fVs[fo + B.cs]; fv2 = B.fVs[fo + 2 * B.cs]; }
		if (ownpml_k) { fi0 = B.fI[fo - fplane]; fi1 = B.fI[fo - fplane + B.cs]; fi2 = B.fI[fo - fplane + 2 * B.cs]; }
		const float4 A = __ldg(p.eA + e), Bc = __ldg(p.eB + e);
		const float4 Ah = __ldg(p.hA + ek_idx), Bh = __ldg(p.hB + ek_idx);
		if (kk + 1 <= e_last) {
			const long long on = o + p.plane, omn = (long long)(kk + 1) * p.plane + rowm;
			n_e = reinterpret_cast<const IdxT*>(p.idx)[on];
			n_h0 = p.Is[on]; n_h1 = p.Is[p.comp + on]; n_h2 = p.Is[2 * p.comp + on];
			n_j0m = p.Is[omn]; n_j2m = p.Is[2 * p.comp + omn];
			n_v0 = p.Vs[on]; n_v1 = p.Vs[p.comp + on]; n_v2 = p.Vs[2 * p.comp + on];
			if (tx == 0 && xc > 0) { n_l1 = p.Is[p.comp + on - 1]; n_l2 = p.Is[2 * p.comp + on - 1]; }
		}
		// ------------------------------------------------------------ E_new(kk)
		{
			const float s1 = __shfl_up_sync(0xffffffffu, h1, 1, 16), s2 = __shfl_up_sync(0xffffffffu, h2, 1, 16);
			if (tx > 0) { l1 = s1; l2 = s2; }
			else if (xc == 0) { l1 = h1; l2 = h2; }
		}
		const bool flag = A.w != 0.0f;
		const bool ownpml = flag && geo && plane_in;
		// cell of another UPML box, or plain cell inside another box's shell footprint: already E_new (that
		// box's k_shell_E ran in place); passed on unchanged
		const bool foreign = (flag && !ownpml) || (!flag && xslab_in_shell(p, xc >> 2, jc, kk));
		const float curl0 = fadd(fsub(fsub(h2, j2m), h1), hk1);
		const float curl1 = fadd(fsub(fsub(h0, hk0), h2), l2);
		const float curl2 = fadd(fsub(fsub(h1, l1), h0), j0m);
		const bool st = cell_ok && own_x && row_in && plane_in && !halo_row && kk < ke;
		if (ownpml) {
			const float4 P0 = __ldg(p.eP0 + e), P1 = __ldg(p.eP1 + e), P2 = __ldg(p.eP2 + e);
			float f0, f1, f2;
			v0 = leap_pml_oop(v0, A.x, Bc.x, curl0, P0.x, P1.x, P2.x, fv0, f0);
			v1 = leap_pml_oop(v1, A.y, Bc.y, curl1, P0.y, P1.y, P2.y, fv1, f1);
			v2 = leap_pml_oop(v2, A.z, Bc.z, curl2, P0.z, P1.z, P2.z, fv2, f2);
			if (st) { B.fVd[fo] = f0; B.fVd[fo + B.cs] = f1; B.fVd[fo + 2 * B.cs] = f2; }
		} else if (!foreign) {
			v0 = leap(v0, A.x, Bc.x, curl0);
			v1 = leap(v1, A.y, Bc.y, curl1);
			v2 = leap(v2, A.z, Bc.z, curl2);
		}
		if (st) { p.Vd[o] = v0; p.Vd[p.comp + o] = v1; p.Vd[2 * p.comp + o] = v2; } // every cell of the slab's chunks
		rV[kk % 3][0][ty][tx] = v0;
		rV[kk % 3][1][ty][tx] = v1;
		rV[kk % 3][2][ty][tx] = v2;
		__syncthreads();

		// ------------------------------------------------------------ H_new(kk-1)
		const int k = kk - 1;
		if (k >= kb && cell_ok && own_x && row_in && plane_k_in && !halo_row) {
			const long long oh = (long long)k * p.plane + row;
			float c0 = hk0, c1 = hk1, c2 = hk2;
			if (k < he && j < p.ny - 1 && x < p.nx - 1 && !foreign_k) { // cells of other boxes: H passed on, their k_shell_H follows
				const float w0 = rV[k % 3][0][ty][tx], w1 = rV[k % 3][1][ty][tx], w2 = rV[k % 3][2][ty][tx];
				const float w0jp = rV[k % 3][0][ty + 1][tx], w2jp = rV[k % 3][2][ty + 1][tx];
				const float w1xp = rV[k % 3][1][ty][tx + 1], w2xp = rV[k % 3][2][ty][tx + 1];
				const float curlh0 = fadd(fsub(fsub(w2, w2jp), w1), v1); // v0, v1: E_new(k+1) of this cell
				const float curlh1 = fadd(fsub(fsub(w0, v0), w2), w2xp);
				const float curlh2 = fadd(fsub(fsub(w1, w1xp), w0), w0jp);
				if (ownpml_k) {
					const float4 P0 = __ldg(p.hP0 + ek_idx), P1 = __ldg(p.hP1 + ek_idx), P2 = __ldg(p.hP2 + ek_idx);
					float* f = B.fI + fo - fplane;
					float f0, f1, f2;
					c0 = leap_pml_oop(c0, Ah.x, Bh.x, curlh0, P0.x, P1.x, P2.x, fi0, f0);
					c1 = leap_pml_oop(c1, Ah.y, Bh.y, curlh1, P0.y, P1.y, P2.y, fi1, f1);
					c2 = leap_pml_oop(c2, Ah.z, Bh.z, curlh2, P0.z, P1.z, P2.z, fi2, f2);
					f[0] = f0; f[B.cs] = f1; f[2 * B.cs] = f2;
				} else {
					c0 = leap(c0, Ah.x, Bh.x, curlh0);
					c1 = leap(c1, Ah.y, Bh.y, curlh1);
					c2 = leap(c2, Ah.z, Bh.z, curlh2);
				}
			}
			if (k < p.kHc1) { p.Id[oh] = c0; p.Id[p.comp + oh] = c1; p.Id[2 * p.comp + oh] = c2; }
		}
		hk0 = h0; hk1 = h1; hk2 = h2;
		ek_idx = e;
		ownpml_k = ownpml; foreign_k = foreign;
	}
	// H of the chunk's last plane when it is not updated here (top of the domain: copy; top plane of a z slab
	// with an upper neighbour: left to the slab-top kernel)
	const int k = e_last;
	if (k == ke - 1 && k >= kb && cell_ok && own_x && row_in && (unsigned)(k - B.s2) < (unsigned)B.n2 && !halo_row && k < p.kHc1) {
		const long long oh = (long long)k * p.plane + row;
		p.Id[oh] = hk0; p.Id[p.comp + oh] = hk1; p.Id[2 * p.comp + oh] = hk2;
	}
}
