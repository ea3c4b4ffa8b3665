// synthetic_operator.cpp -- see synthetic_operator.h.
// Host-only C++ (no CUDA): the operator build stays on the host like the reference's
// Operator::CalcECOperator; only the result is uploaded.
#include <cuda_runtime.h>
#include "synthetic_operator.h"

#include <omp.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../entry_set.h"

namespace {

constexpr double EPS0 = 8.85418781762e-12; // tools/constants.h:23-26
constexpr double MUE0 = 1.256637062e-6;
constexpr double C0 = 299792458.0;
constexpr double Z0 = 376.730313461;
constexpr double PI = 3.141592653589793238462643383279;
constexpr int MAX_ORDER = 8;

enum { P_MATERIAL = 0, P_METAL, P_LORENTZ, P_EXCITATION };
constexpr unsigned MASK_MAT = (1u << P_MATERIAL) | (1u << P_LORENTZ);
constexpr unsigned MASK_MAT_METAL = MASK_MAT | (1u << P_METAL);
constexpr unsigned MASK_EXC = 1u << P_EXCITATION;

struct Prop {
	int type = 0, prio = 0;
	double start[3], stop[3];
	double epsR = 1, mueR = 1, kappa = 0, sigma = 0;
	int order = 0;
	double eps_fp[MAX_ORDER] = {0}, eps_tau[MAX_ORDER] = {0}, eps_flor[MAX_ORDER] = {0};
	double mue_fp[MAX_ORDER] = {0}, mue_tau[MAX_ORDER] = {0}, mue_flor[MAX_ORDER] = {0};
	int exc_type = 0;
	double exc_vec[3] = {0, 0, 0}, delay = 0;
};

struct MurPlaneH {
	int ny; unsigned line, shift, n[2], start_ts;
	std::vector<float> cP, cPP;
};
struct UpmlBoxH { unsigned start[3], n[3]; };
struct LorOrderH {
	bool volt_on = false, curr_on = false, volt_lor_on = false, curr_lor_on = false;
	std::vector<unsigned> pos[3];
	std::vector<float> c[6][3]; // v_int v_ext v_lor i_int i_ext i_lor
};

struct PlaneEC { std::vector<float> C[3], G[3], L[3], R[3]; };

// everything of a cell that later passes need besides the table entry
struct PlaneOp {
	std::vector<uint32_t> index; // [ny][nx] into the global table
};

} // namespace

struct oems_synth {
	unsigned N[3];
	std::vector<double> Ls[3];
	double gd = 1;
	int bc[6] = {0, 0, 0, 0, 0, 0};
	unsigned pml[6] = {8, 8, 8, 8, 8, 8};
	double bg[4] = {1, 1, 0, 0}; // epsR mueR kappa sigma
	double forced_dT = 0, factor = 1;
	std::vector<Prop> props;
	int exc_kind = -1;
	double exc_f0 = 0, exc_fc = 0, exc_period = 0;
	std::string err;

	// results
	bool built = false;
	bool pinned = false; // index registered as page-locked memory (oems_synth_pin)
	double dT = 0;
	unsigned nyquist = 0, sig_len = 0;
	std::vector<float> sig[2];
	EntrySet table;
	std::vector<uint16_t> idx16;
	std::vector<uint32_t> idx32;
	// the same index as unique xy planes + one plane id per z (what oems_synth_upload sends)
	std::vector<unsigned> plane_of_z;
	std::vector<uint16_t> uplanes16;
	std::vector<uint32_t> uplanes32;
	int index_bytes = 0;
	unsigned unique_planes = 0;
	// z-slab restricted build (one process per GPU): only the planes [slab_lo, slab_hi) this rank holds are built
	bool slab_set = false;
	unsigned slab_lo = 0, slab_hi = 0;
	bool held(unsigned k) const { return !slab_set || (k >= slab_lo && k < slab_hi); }
	uint32_t cell_id(unsigned k, size_t p2) const
	{
		const size_t np = (size_t)N[0] * N[1];
		return index_bytes == 2 ? uplanes16[(size_t)plane_of_z[k] * np + p2] : uplanes32[(size_t)plane_of_z[k] * np + p2];
	}
	void ensure_full_index()
	{ // the per-cell index [nz][ny][nx]; only materialised for callers that ask for it (oems_synth_index / pin)
		const size_t np = (size_t)N[0] * N[1];
		const long long Nz = N[2];
		if (index_bytes == 2 && idx16.empty()) {
			idx16.resize(np * Nz);
#pragma omp parallel for schedule(static)
			for (long long k = 0; k < Nz; ++k) memcpy(idx16.data() + (size_t)k * np, uplanes16.data() + (size_t)plane_of_z[k] * np, np * 2);
		} else if (index_bytes == 4 && idx32.empty()) {
			idx32.resize(np * Nz);
#pragma omp parallel for schedule(static)
			for (long long k = 0; k < Nz; ++k) memcpy(idx32.data() + (size_t)k * np, uplanes32.data() + (size_t)plane_of_z[k] * np, np * 4);
		}
	}
	std::vector<unsigned> exc_idx[2][3], exc_dir[2], exc_delay[2];
	std::vector<float> exc_amp[2];
	std::vector<MurPlaneH> mur;
	std::vector<UpmlBoxH> upml;
	int lor_order = 0;
	std::vector<LorOrderH> lor;

	// ---- mesh helpers: Operator::GetDiscLine/GetDiscDelta/GetEdgeLength/GetNodeWidth/GetNodeArea
	// (FDTD/operator.cpp:143-240, operator.h:174,194)
	double disc_line(int n, unsigned pos, bool dual) const
	{
		if (pos >= N[n]) return 0.0;
		const double* L = Ls[n].data();
		if (!dual) return L[pos];
		if (pos < N[n] - 1) return 0.5 * (L[pos] + L[pos + 1]);
		return L[pos] + 0.5 * (L[pos] - L[pos - 1]);
	}
	double disc_delta(int n, unsigned pos, bool dual) const
	{
		if (pos >= N[n]) return 0.0;
		if (!dual) {
			if (pos < N[n] - 1) return disc_line(n, pos + 1, false) - disc_line(n, pos, false);
			return disc_line(n, pos, false) - disc_line(n, pos - 1, false);
		}
		if (pos > 0) return disc_line(n, pos, true) - disc_line(n, pos - 1, true);
		return disc_line(n, 1, false) - disc_line(n, 0, false);
	}
	double edge_length(int n, const unsigned pos[3], bool dual) const { return disc_delta(n, pos[n], dual) * gd; }
	double node_width(int n, const unsigned pos[3], bool dual) const { return edge_length(n, pos, !dual); }
	double node_area(int n, const unsigned pos[3], bool dual) const
	{
		return node_width((n + 1) % 3, pos, dual) * node_width((n + 2) % 3, pos, dual);
	}
	bool yee_coords(int ny, const unsigned pos[3], double* c, bool dual) const
	{
		for (int n = 0; n < 3; ++n) c[n] = disc_line(n, pos[n], dual);
		c[ny] = disc_line(ny, pos[ny], !dual);
		if (!dual) return pos[ny] < N[ny] - 1;
		const int nP = (ny + 1) % 3, nPP = (ny + 2) % 3;
		return !(pos[nP] >= N[nP] - 1 || pos[nPP] >= N[nPP] - 1);
	}

	// stand-in for CSXCAD's GetPropertyByCoordPriority: highest priority box containing the
	// point, ties to the box added later
	const Prop* prop_at(const double c[3], unsigned mask) const
	{
		const Prop* best = nullptr;
		for (const Prop& q : props) {
			if (!(mask & (1u << q.type))) continue;
			if (c[0] < q.start[0] || c[0] > q.stop[0] || c[1] < q.start[1] || c[1] > q.stop[1] || c[2] < q.start[2] || c[2] > q.stop[2])
				continue;
			if (!best || q.prio >= best->prio) best = &q;
		}
		return best;
	}
	double material(const double c[3], int type) const
	{ // Operator::GetMaterial operator.cpp:1289-1345
		const Prop* p = prop_at(c, MASK_MAT);
		if (p) return type == 0 ? p->epsR : type == 1 ? p->kappa : type == 2 ? p->mueR : p->sigma;
		return type == 0 ? bg[0] : type == 1 ? bg[2] : type == 2 ? bg[1] : bg[3];
	}
	bool cell_center(const int pos[3], double c[3]) const
	{
		for (int n = 0; n < 3; ++n)
			if (pos[n] < 0 || pos[n] >= (int)N[n]) return false;
		for (int n = 0; n < 3; ++n) c[n] = disc_line(n, (unsigned)pos[n], true);
		return true;
	}
	// Operator::AverageMatCellCenter operator.cpp:1347-1444
	void eff_mat(int ny, const unsigned pos[3], double E[4]) const
	{
		const int n = ny, nP = (n + 1) % 3, nPP = (n + 2) % 3;
		int lp[3] = {(int)pos[0], (int)pos[1], (int)pos[2]};
		double c[3], area = 0;
		E[0] = E[1] = E[2] = E[3] = 0;
		auto acc_eps = [&]() {
			if (cell_center(lp, c)) {
				const unsigned up[3] = {(unsigned)lp[0], (unsigned)lp[1], (unsigned)lp[2]};
				const double A = node_area(ny, up, true);
				E[0] += material(c, 0) * A;
				E[1] += material(c, 1) * A;
				area += A;
			}
		};
		acc_eps();
		--lp[nP]; acc_eps();
		++lp[nP]; --lp[nPP]; acc_eps();
		--lp[nP]; acc_eps();
		E[0] *= EPS0 / area;
		E[1] /= area;
		lp[0] = pos[0]; lp[1] = pos[1]; lp[2] = pos[2];
		double length = 0;
		auto acc_mue = [&]() {
			if (cell_center(lp, c)) {
				const unsigned up[3] = {(unsigned)lp[0], (unsigned)lp[1], (unsigned)lp[2]};
				const double d = node_width(n, up, true);
				E[2] += d / material(c, 2);
				const double sig = material(c, 3);
				if (sig) E[3] += d / sig; else E[3] = 0;
				length += d;
			}
		};
		--lp[n]; acc_mue();
		++lp[n]; acc_mue();
		E[2] = length * MUE0 / E[2];
		if (E[3]) E[3] = length / E[3];
	}
	// Operator::Calc_ECPos operator.cpp:1189-1256
	void calc_ec_pos(int ny, const unsigned pos[3], double EC[4]) const
	{
		double E[4];
		eff_mat(ny, pos, E);
		double delta = edge_length(ny, pos, false), area = node_area(ny, pos, false);
		if (delta) { EC[0] = E[0] * area / delta; EC[1] = E[1] * area / delta; } else { EC[0] = EC[1] = 0; }
		delta = edge_length(ny, pos, true);
		area = node_area(ny, pos, true);
		if (delta) { EC[2] = E[2] * area / delta; EC[3] = E[3] * area / delta; } else { EC[2] = EC[3] = 0; }
	}

	void ec_plane(unsigned k, PlaneEC& out) const
	{
		const size_t np = (size_t)N[0] * N[1];
		for (int n = 0; n < 3; ++n) { out.C[n].resize(np); out.G[n].resize(np); out.L[n].resize(np); out.R[n].resize(np); }
#pragma omp parallel for schedule(static)
		for (long long j = 0; j < (long long)N[1]; ++j)
			for (unsigned i = 0; i < N[0]; ++i) {
				const unsigned pos[3] = {i, (unsigned)j, k};
				const size_t p = (size_t)j * N[0] + i;
				for (int n = 0; n < 3; ++n) {
					double EC[4];
					calc_ec_pos(n, pos, EC);
					out.C[n][p] = (float)EC[0]; out.G[n][p] = (float)EC[1];
					out.L[n][p] = (float)EC[2]; out.R[n][p] = (float)EC[3];
				}
			}
	}

	static int refl(int q, int Nn)
	{ // AdrOp::GetPos with SetReflection2Cell tools/AdrOp.cpp:183-222
		if (q < 0) q = -q - 1;
		if (q > Nn - 1) q = 2 * (Nn - 1) - q + 1;
		return q;
	}

	// Operator::CalcTimestep_Var3 operator.cpp:1956-2030 restricted to plane k; E[0..2] are
	// the EC planes refl(k-1), k, refl(k+1)
	double timestep_plane(unsigned k, const PlaneEC* const E[3]) const
	{
		const int Nx = (int)N[0], Ny = (int)N[1], Nz = (int)N[2];
		(void)Nz;
		// z offsets are only ever -1, 0, +1: E[d2+1] already holds the reflected plane
		double dTmin = 1e200;
#pragma omp parallel for schedule(static) reduction(min : dTmin)
		for (int j = 0; j < Ny; ++j)
			for (int i = 0; i < Nx; ++i)
				for (int n = 0; n < 3; ++n) {
					const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
					const int pos[3] = {i, j, (int)k};
					// EC_L / EC_C are FDTD_FLOAT arrays (operator.h:346-349): each product L*C, each 1/(L*C) and each
					// two-term sum on the right-hand sides of operator.cpp:1984-1990 is a FLOAT expression in the
					// reference and only its result is widened to double -- keep `at` float and the casts below
					auto at = [&](int comp, bool isL, int d0, int d1, int d2) -> float {
						const int q[2] = {refl(pos[0] + d0, Nx), refl(pos[1] + d1, Ny)};
						const PlaneEC& P = *E[d2 + 1];
						const size_t p = (size_t)q[1] * Nx + q[0];
						return isL ? P.L[comp][p] : P.C[comp][p];
					};
					auto sh = [&](int a, int sa, int b, int sb, int out[3]) {
						out[0] = out[1] = out[2] = 0;
						out[a] += sa; out[b] += sb;
					};
					int d[3], d1[3], d2[3], dp[3];
					// wqp
					sh(n, 0, n, 0, d);
					const float LPP0 = at(nPP, true, 0, 0, 0), LP0 = at(nP, true, 0, 0, 0), Cn0 = at(n, false, 0, 0, 0);
					sh(nP, 1, n, 0, d);
					double wqp = (float)(1 / (LPP0 * at(n, false, d[0], d[1], d[2])) + 1 / (LPP0 * Cn0));
					sh(nPP, 1, n, 0, d);
					wqp += (float)(1 / (LP0 * at(n, false, d[0], d[1], d[2])) + 1 / (LP0 * Cn0));
					sh(nP, -1, n, 0, d1);
					const float LPP1 = at(nPP, true, d1[0], d1[1], d1[2]), Cn1 = at(n, false, d1[0], d1[1], d1[2]);
					wqp += (float)(1 / (LPP1 * Cn0) + 1 / (LPP1 * Cn1));
					sh(nP, -1, nPP, -1, d2);
					const float LP2 = at(nP, true, d2[0], d2[1], d2[2]), Cn2 = at(n, false, d2[0], d2[1], d2[2]);
					wqp += (float)(1 / (LP2 * Cn1) + 1 / (LP2 * Cn2));
					// wt1
					const float CP0 = at(nP, false, 0, 0, 0), CPP0 = at(nPP, false, 0, 0, 0);
					sh(nPP, -1, n, 0, dp);
					const float LPPm = LPP1; // L[nPP] at pos - nP
					const float LPm = at(nP, true, dp[0], dp[1], dp[2]); // L[nP] at pos - nPP
					double w4[4] = {(float)(1 / (LPP0 * CP0)), (float)(1 / (LPPm * CP0)), (float)(1 / (LP0 * CPP0)), (float)(1 / (LPm * CPP0))};
					double mn = w4[0]; // min() of operator.cpp:1942-1952: plain '<' scan
					for (int q = 1; q < 4; ++q) if (w4[q] < mn) mn = w4[q];
					const double wt1 = w4[0] + w4[1] + w4[2] + w4[3] - 2 * mn;
					// wt2
					sh(n, 1, n, 0, d);
					const float CPn = at(nP, false, d[0], d[1], d[2]), CPPn = at(nPP, false, d[0], d[1], d[2]);
					double v4[4] = {(float)(1 / (LPP0 * CPn)), (float)(1 / (LPPm * CPn)), (float)(1 / (LP0 * CPPn)), (float)(1 / (LPm * CPPn))};
					mn = v4[0];
					for (int q = 1; q < 4; ++q) if (v4[q] < mn) mn = v4[q];
					const double wt2 = v4[0] + v4[1] + v4[2] + v4[3] - 2 * mn;
					const double w_total = wqp + wt1 + wt2;
					const double newT = 2 / std::sqrt(w_total);
					if (newT < dTmin && newT > 0.0) dTmin = newT;
				}
		return dTmin;
	}

	// ---- UPML grading: Operator_Ext_UPML::CalcGradingKappa operator_ext_upml.cpp:269-343 with the
	// default grading function (:30) evaluated directly (fparser is not vendored)
	static double grading(double D, double dl, double W, double Z)
	{
		return -std::log(1e-6) * std::log(2.5) / (2 * dl * Z * (std::pow(2.5, W / dl) - 1)) * std::pow(2.5, D / dl);
	}
	void grading_kappa(int ny, const unsigned pos[3], double kv[3], double ki[3]) const
	{
		double depth = 0, width = 0;
		for (int n = 0; n < 3; ++n) {
			const unsigned Nn = N[n];
			if (pos[n] <= pml[2 * n] && bc[2 * n] == 3) {
				width = (disc_line(n, pml[2 * n], false) - disc_line(n, 0, false)) * gd;
				depth = width - (disc_line(n, pos[n], false) - disc_line(n, 0, false)) * gd;
				if (n == ny) depth -= edge_length(n, pos, false) / 2;
				const double dl = width / pml[2 * n];
				kv[n] = depth > 0 ? grading(depth, dl, width, Z0) : 0;
				if (n == ny) depth += edge_length(n, pos, false) / 2;
				if (n != ny) depth -= edge_length(n, pos, false) / 2;
				if (depth < 0) depth = 0;
				ki[n] = depth > 0 ? grading(depth, dl, width, Z0) : 0;
			} else if (pos[n] >= Nn - 1 - pml[2 * n + 1] && bc[2 * n + 1] == 3) {
				width = (disc_line(n, Nn - 1, false) - disc_line(n, Nn - pml[2 * n + 1] - 1, false)) * gd;
				depth = width - (disc_line(n, Nn - 1, false) - disc_line(n, pos[n], false)) * gd;
				if (n == ny) depth += edge_length(n, pos, false) / 2;
				const double dl = width / pml[2 * n]; // quirk kept: lower side's size (:319)
				kv[n] = depth > 0 ? grading(depth, dl, width, Z0) : 0;
				if (n == ny) depth -= edge_length(n, pos, false) / 2;
				if (n != ny) depth += edge_length(n, pos, false) / 2;
				if (depth > width) depth = 0;
				ki[n] = depth > 0 ? grading(depth, dl, width, Z0) : 0;
			} else { kv[n] = 0; ki[n] = 0; }
		}
	}
	bool in_upml(const unsigned pos[3]) const
	{
		for (const UpmlBoxH& B : upml) {
			bool in = true;
			for (int a = 0; a < 3; ++a) in &= pos[a] - B.start[a] < B.n[a];
			if (in) return true;
		}
		return false;
	}

	// coefficient tuple of one cell: Calc_ECOperatorPos (operator.cpp:956-984), ApplyElectricBC
	// (:1099-1138), CalcPEC_Range (:2046-2084), ApplyMagneticBC (:1140-1187), hard-source zeroing
	// (operator_ext_excitation.cpp:186-190,218-222), UPML (operator_ext_upml.cpp:345-445)
	void cell_entry(const unsigned pos[3], const PlaneEC& E, bool any_metal, bool any_exc, oems_coeff_entry& e) const
	{
		memset(&e, 0, sizeof(e));
		const size_t p = (size_t)pos[1] * N[0] + pos[0];
		for (int n = 0; n < 3; ++n) {
			const double C = E.C[n][p], G = E.G[n][p];
			if (C > 0) {
				e.vv[n] = (float)((1.0 - dT * G / 2.0 / C) / (1.0 + dT * G / 2.0 / C));
				e.vi[n] = (float)((dT / C) / (1.0 + dT * G / 2.0 / C));
			}
			const double L = E.L[n][p], R = E.R[n][p];
			if (L > 0) {
				e.ii[n] = (float)((1.0 - dT * R / 2.0 / L) / (1.0 + dT * R / 2.0 / L));
				e.iv[n] = (float)((dT / L) / (1.0 + dT * R / 2.0 / L));
			}
		}
		for (int n = 0; n < 3; ++n) {
			const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
			if (bc[2 * n] != -1 && pos[n] == 0) { e.vv[nP] = e.vi[nP] = 0; e.vv[nPP] = e.vi[nPP] = 0; }
			if (bc[2 * n + 1] != -1 && pos[n] == N[n] - 1)
				for (int c = 0; c < 3; ++c) e.vv[c] = e.vi[c] = 0;
		}
		double c[3];
		if (any_metal)
			for (int n = 0; n < 3; ++n) {
				yee_coords(n, pos, c, false);
				const Prop* q = prop_at(c, MASK_MAT_METAL);
				if (q && q->type == P_METAL) e.vv[n] = e.vi[n] = 0;
			}
		for (int n = 0; n < 3; ++n) {
			const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
			if (bc[2 * n] == 1 && pos[n] == 0)
				for (int cc = 0; cc < 3; ++cc) e.ii[cc] = e.iv[cc] = 0;
			if (bc[2 * n + 1] == 1 && pos[n] == N[n] - 2) { e.ii[nP] = e.iv[nP] = 0; e.ii[nPP] = e.iv[nPP] = 0; }
			if (pos[n] == N[n] - 1)
				for (int cc = 0; cc < 3; ++cc) e.ii[cc] = e.iv[cc] = 0;
		}
		if (any_exc) {
			for (int n = 0; n < 3; ++n) {
				if (!yee_coords(n, pos, c, false)) continue;
				const Prop* q = prop_at(c, MASK_EXC);
				if (q && q->exc_type == 1) e.vv[n] = e.vi[n] = 0;   // ActiveDir is true for all components (CSPropExcitation default)
			}
			for (int n = 0; n < 3; ++n) {
				if (pos[0] >= N[0] - 1 || pos[1] >= N[1] - 1 || pos[2] >= N[2] - 1) continue;
				if (!yee_coords(n, pos, c, true)) continue;
				const Prop* q = prop_at(c, MASK_EXC);
				if (q && q->exc_type == 3) e.ii[n] = e.iv[n] = 0;
			}
		}
		if (!upml.empty() && in_upml(pos)) {
			e.pml = 1.0f;
			for (int n = 0; n < 3; ++n) {
				double em[4], kv[3] = {0, 0, 0}, ki[3] = {0, 0, 0};
				eff_mat(n, pos, em);
				grading_kappa(n, pos, kv, ki);
				const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
				if ((kv[0] + kv[1] + kv[2]) != 0 && em[1] < 1e3) {
					if ((e.vv[n] + e.vi[n]) != 0) {
						e.vv[n] = (float)((2 * EPS0 - kv[nP] * dT) / (2 * EPS0 + kv[nP] * dT));
						e.vi[n] = (float)((2 * EPS0 * dT) / (2 * EPS0 + kv[nP] * dT) * edge_length(n, pos, false) / node_area(n, pos, false));
						e.pml_vv[n] = (float)((2 * EPS0 - kv[nPP] * dT) / (2 * EPS0 + kv[nPP] * dT));
						e.pml_vvfn[n] = (float)((2 * EPS0 + kv[n] * dT) / (2 * EPS0 + kv[nPP] * dT) / em[0]);
						e.pml_vvfo[n] = (float)((2 * EPS0 - kv[n] * dT) / (2 * EPS0 + kv[nPP] * dT) / em[0]);
					}
				} else {
					e.pml_vv[n] = e.vv[n];
					e.vv[n] = 0;
					e.pml_vvfo[n] = 0;
					e.pml_vvfn[n] = 1;
				}
				if ((ki[0] + ki[1] + ki[2]) != 0) {
					if ((e.ii[n] + e.iv[n]) != 0) {
						e.ii[n] = (float)((2 * EPS0 - ki[nP] * dT) / (2 * EPS0 + ki[nP] * dT));
						e.iv[n] = (float)((2 * EPS0 * dT) / (2 * EPS0 + ki[nP] * dT) * edge_length(n, pos, true) / node_area(n, pos, true));
						e.pml_ii[n] = (float)((2 * EPS0 - ki[nPP] * dT) / (2 * EPS0 + ki[nPP] * dT));
						e.pml_iifn[n] = (float)((2 * EPS0 + ki[n] * dT) / (2 * EPS0 + ki[nPP] * dT) / em[2]);
						e.pml_iifo[n] = (float)((2 * EPS0 - ki[n] * dT) / (2 * EPS0 + ki[nPP] * dT) / em[2]);
					}
				} else {
					e.pml_ii[n] = e.ii[n];
					e.ii[n] = 0;
					e.pml_iifo[n] = 0;
					e.pml_iifn[n] = 1;
				}
			}
		}
	}

	// z-signature of plane k: everything z-dependent that enters planes k-1..k+1 (EC, timestep,
	// coefficients, PML).  Planes with equal signatures are computed once.
	std::vector<uint64_t> zsig(unsigned k) const
	{
		std::vector<uint64_t> s;
		const int Nz = (int)N[2];
		auto bits = [](double v) { uint64_t u; memcpy(&u, &v, 8); return u; };
		for (int m = -2; m <= 2; ++m) {
			const int q = (int)k + m;
			if (q < 0 || q >= Nz) { s.push_back(~0ull); s.push_back(0); s.push_back(0); s.push_back(0); continue; }
			s.push_back(bits(disc_delta(2, (unsigned)q, false)));
			s.push_back(bits(disc_delta(2, (unsigned)q, true)));
			uint64_t line_mask = 0, mid_mask = 0;
			const double zl = disc_line(2, (unsigned)q, false), zm = disc_line(2, (unsigned)q, true);
			for (size_t p = 0; p < props.size() && p < 64; ++p) {
				if (zl >= props[p].start[2] && zl <= props[p].stop[2]) line_mask |= 1ull << p;
				if (zm >= props[p].start[2] && zm <= props[p].stop[2]) mid_mask |= 1ull << p;
			}
			s.push_back(line_mask);
			s.push_back(mid_mask);
		}
		// distance classes to the faces (BC lines, PML depth); 3 = "far"
		const unsigned lo_span = std::max(3u, (bc[4] == 3 ? pml[4] : 0) + 3), hi_span = std::max(3u, (bc[5] == 3 ? pml[5] : 0) + 3);
		s.push_back(k < lo_span ? k : lo_span);
		s.push_back((unsigned)(Nz - 1) - k < hi_span ? (unsigned)(Nz - 1) - k : hi_span);
		// UPML box membership pattern in z
		uint64_t boxmask = 0;
		for (size_t b = 0; b < upml.size(); ++b)
			if (k - upml[b].start[2] < upml[b].n[2]) boxmask |= 1ull << b;
		s.push_back(boxmask);
		return s;
	}
};

// ======================================================================== C interface
extern "C" {

oems_synth* oems_synth_create(unsigned nx, unsigned ny, unsigned nz, const double* x, const double* y, const double* z, double grid_delta)
{
	if (nx < 3 || ny < 3 || nz < 3 || !x || !y || !z) return nullptr;
	oems_synth* s = new oems_synth;
	s->N[0] = nx; s->N[1] = ny; s->N[2] = nz;
	s->Ls[0].assign(x, x + nx); s->Ls[1].assign(y, y + ny); s->Ls[2].assign(z, z + nz);
	s->gd = grid_delta;
	return s;
}
void oems_synth_destroy(oems_synth* s)
{
	if (s && s->pinned) cudaHostUnregister(const_cast<void*>(oems_synth_index(s)));
	delete s;
}
void oems_synth_set_bc(oems_synth* s, const int bc[6], const unsigned pml_size[6])
{
	for (int n = 0; n < 6; ++n) { s->bc[n] = bc[n]; if (pml_size) s->pml[n] = pml_size[n]; }
}
void oems_synth_set_background(oems_synth* s, double epsR, double mueR, double kappa, double sigma)
{
	s->bg[0] = epsR; s->bg[1] = mueR; s->bg[2] = kappa; s->bg[3] = sigma;
}
void oems_synth_set_timestep(oems_synth* s, double forced_dT, double factor)
{
	s->forced_dT = forced_dT; s->factor = factor > 0 ? factor : 1.0;
}
static Prop& new_prop(oems_synth* s, int type, int prio, const double a[3], const double b[3])
{
	Prop p;
	p.type = type; p.prio = prio;
	for (int n = 0; n < 3; ++n) { p.start[n] = std::min(a[n], b[n]); p.stop[n] = std::max(a[n], b[n]); }
	s->props.push_back(p);
	return s->props.back();
}
int oems_synth_add_material(oems_synth* s, int prio, const double a[3], const double b[3], double epsR, double mueR, double kappa, double sigma)
{
	if (s->props.size() >= 64) return -1;
	Prop& p = new_prop(s, P_MATERIAL, prio, a, b);
	p.epsR = epsR; p.mueR = mueR; p.kappa = kappa; p.sigma = sigma;
	return (int)s->props.size() - 1;
}
int oems_synth_add_metal(oems_synth* s, int prio, const double a[3], const double b[3])
{
	if (s->props.size() >= 64) return -1;
	new_prop(s, P_METAL, prio, a, b);
	return (int)s->props.size() - 1;
}
int oems_synth_add_lorentz(oems_synth* s, int prio, const double a[3], const double b[3], double epsR, double mueR, double kappa,
                           double sigma, int order, const double* eps_fp, const double* eps_tau, const double* eps_flor,
                           const double* mue_fp, const double* mue_tau, const double* mue_flor)
{
	if (s->props.size() >= 64 || order > MAX_ORDER) return -1;
	Prop& p = new_prop(s, P_LORENTZ, prio, a, b);
	p.epsR = epsR; p.mueR = mueR; p.kappa = kappa; p.sigma = sigma; p.order = order;
	for (int o = 0; o < order; ++o) {
		p.eps_fp[o] = eps_fp ? eps_fp[o] : 0; p.eps_tau[o] = eps_tau ? eps_tau[o] : 0; p.eps_flor[o] = eps_flor ? eps_flor[o] : 0;
		p.mue_fp[o] = mue_fp ? mue_fp[o] : 0; p.mue_tau[o] = mue_tau ? mue_tau[o] : 0; p.mue_flor[o] = mue_flor ? mue_flor[o] : 0;
	}
	return (int)s->props.size() - 1;
}
int oems_synth_add_excitation(oems_synth* s, int prio, const double a[3], const double b[3], int exc_type, const double vec[3], double delay_s)
{
	if (s->props.size() >= 64) return -1;
	Prop& p = new_prop(s, P_EXCITATION, prio, a, b);
	p.exc_type = exc_type;
	for (int n = 0; n < 3; ++n) p.exc_vec[n] = vec[n];
	p.delay = delay_s;
	return (int)s->props.size() - 1;
}
void oems_synth_set_excite_gauss(oems_synth* s, double f0, double fc) { s->exc_kind = 0; s->exc_f0 = f0; s->exc_fc = fc; s->exc_period = 0; }
void oems_synth_set_excite_sinus(oems_synth* s, double f0) { s->exc_kind = 1; s->exc_f0 = f0; s->exc_period = 1 / f0; }
const char* oems_synth_last_error(const oems_synth* s) { return s ? s->err.c_str() : "null builder"; }

static unsigned calc_nyquist(double fmax, double dT)
{ // tools/useful.cpp:30-36
	if (fmax == 0) return UINT_MAX;
	if (dT == 0) return 1;
	return (unsigned)std::floor(1 / fmax / 2 / dT);
}

int oems_synth_set_slab(oems_synth* s, unsigned z_begin, unsigned z_end)
{
	if (!s || s->built) return 1;
	if (z_begin >= z_end || z_end > s->N[2]) { s->err = "set_slab: bad range"; return 1; }
	s->slab_set = true;
	s->slab_lo = z_begin > 0 ? z_begin - 1 : 0;            // plus the ghost planes the slab engine holds
	s->slab_hi = z_end < s->N[2] ? z_end + 1 : s->N[2];
	return 0;
}

// Operator::CalcTimestep over the planes of this rank's slab only (all planes without a slab): the ranks take the
// minimum of these values (one MPI/NCCL/gloo MIN-reduction of a double) and pass it to oems_synth_set_timestep
int oems_synth_local_timestep(oems_synth* s, double* dT_out)
{
	if (!s || !dT_out) return 1;
	const unsigned Nz = s->N[2];
	std::map<std::vector<uint64_t>, unsigned> seen;
	double dT = 1e200;
	std::map<unsigned, PlaneEC> cache;
	for (unsigned k = 0; k < Nz; ++k) {
		if (!s->held(k)) continue;
		if (!seen.emplace(s->zsig(k), k).second) continue;
		const int km = oems_synth::refl((int)k - 1, (int)Nz), kp = oems_synth::refl((int)k + 1, (int)Nz);
		for (int q : {km, (int)k, kp})
			if (!cache.count((unsigned)q)) s->ec_plane((unsigned)q, cache[(unsigned)q]);
		const PlaneEC* E[3] = {&cache[(unsigned)km], &cache[k], &cache[(unsigned)kp]};
		dT = std::min(dT, s->timestep_plane(k, E));
		for (auto it = cache.begin(); it != cache.end();)
			it = (it->first + 1 < k) ? cache.erase(it) : std::next(it);
	}
	*dT_out = dT;
	return 0;
}

int oems_synth_build(oems_synth* s, unsigned max_ts)
{
	if (s->built) { s->err = "already built"; return 1; }
	if (s->exc_kind < 0) { s->err = "no excitation signal set"; return 1; }
	const unsigned Nx = s->N[0], Ny = s->N[1], Nz = s->N[2];
	bool any_metal = false, any_exc = false, any_lor = false;
	for (const Prop& p : s->props) { any_metal |= p.type == P_METAL; any_exc |= p.type == P_EXCITATION; any_lor |= p.type == P_LORENTZ; }

	// ---- UPML boxes: Operator_Ext_UPML::Create_UPML operator_ext_upml.cpp:69-247
	{
		int BC[6]; unsigned size[6];
		for (int n = 0; n < 6; ++n) { BC[n] = s->bc[n]; size[n] = s->pml[n]; }
		for (int n = 0; n < 3; ++n)
			if ((size[2 * n] * (BC[2 * n] == 3) + size[2 * n + 1] * (BC[2 * n + 1] == 3)) >= s->N[n]) {
				BC[2 * n] = 0; size[2 * n] = 0; BC[2 * n + 1] = 0; size[2 * n + 1] = 0;
			}
		for (int n = 0; n < 6; ++n) { s->bc[n] = BC[n] == 3 ? 3 : (s->bc[n] == 3 ? 0 : s->bc[n]); s->pml[n] = size[n]; }
		unsigned start[3] = {0, 0, 0}, stop[3] = {Nx - 1, Ny - 1, Nz - 1};
		auto add = [&]() {
			UpmlBoxH B;
			for (int q = 0; q < 3; ++q) { B.start[q] = start[q]; B.n[q] = stop[q] - start[q] + 1; }
			s->upml.push_back(B);
		};
		if (BC[0] == 3) { start[0] = 0; stop[0] = size[0]; add(); }
		if (BC[1] == 3) { start[0] = Nx - 1 - size[1]; stop[0] = Nx - 1; add(); }
		start[0] = (size[0] + 1) * (BC[0] == 3);
		stop[0] = Nx - 1 - (size[0] + 1) * (BC[1] == 3); // size[0]: as in the reference (:164)
		if (BC[2] == 3) { start[1] = 0; stop[1] = size[2]; add(); }
		if (BC[3] == 3) { start[1] = Ny - 1 - size[3]; stop[1] = Ny - 1; add(); }
		start[1] = (size[2] + 1) * (BC[2] == 3);
		stop[1] = Ny - 1 - (size[3] + 1) * (BC[3] == 3);
		if (BC[4] == 3) { start[2] = 0; stop[2] = size[4]; add(); }
		if (BC[5] == 3) { start[2] = Nz - 1 - size[5]; stop[2] = Nz - 1; add(); }
	}

	// ---- group planes by z-signature (with a slab set: only the planes this rank holds; the others map to class 0
	// and are never looked at by the slab engine)
	std::map<std::vector<uint64_t>, unsigned> sig2id;
	std::vector<unsigned> plane_id(Nz, 0), rep; // representative plane of every class
	for (unsigned k = 0; k < Nz; ++k) {
		if (!s->held(k)) continue;
		auto it = sig2id.find(s->zsig(k));
		if (it == sig2id.end()) { it = sig2id.emplace(s->zsig(k), (unsigned)rep.size()).first; rep.push_back(k); }
		plane_id[k] = it->second;
	}
	s->unique_planes = (unsigned)rep.size();

	// ---- timestep (operator.cpp:994-1025): minimum over the representative planes (of this rank's slab: ranks agree
	// on the global minimum through oems_synth_local_timestep + set_timestep before they build)
	if (s->forced_dT > 0) s->dT = s->forced_dT;
	else {
		double dT = 1e200;
		std::map<unsigned, PlaneEC> cache; // EC planes, a few alive at a time
		for (unsigned k : rep) {
			const int km = oems_synth::refl((int)k - 1, (int)Nz), kp = oems_synth::refl((int)k + 1, (int)Nz);
			for (int q : {km, (int)k, kp})
				if (!cache.count((unsigned)q)) s->ec_plane((unsigned)q, cache[(unsigned)q]);
			const PlaneEC* E[3] = {&cache[(unsigned)km], &cache[k], &cache[(unsigned)kp]};
			dT = std::min(dT, s->timestep_plane(k, E));
			for (auto it = cache.begin(); it != cache.end();)
				it = (it->first + 1 < k) ? cache.erase(it) : std::next(it);
		}
		s->dT = dT;
	}
	s->dT *= s->factor;
	if (s->exc_period > 0) {
		const unsigned TS = (unsigned)std::ceil(s->exc_period / s->dT);
		s->dT = s->exc_period / TS;
	}
	const double dT = s->dT;

	// ---- excitation signal: Excitation::CalcGaussianPulsExcitation / CalcSinusExcitation
	// FDTD/excitation.cpp:150-176,254-276
	if (s->exc_kind == 0) {
		unsigned len = (unsigned)std::ceil(2.0 * 9.0 / (2.0 * PI * s->exc_fc) / dT);
		if (len > max_ts) len = max_ts;
		s->sig_len = len;
		s->sig[0].assign(len, 0.f); s->sig[1].assign(len, 0.f);
		const double f0 = s->exc_f0, fc = s->exc_fc;
		for (unsigned n = 0; n < len; ++n) {
			double t = n * dT;
			s->sig[0][n] = (float)(std::cos(2.0 * PI * f0 * (t - 9.0 / (2.0 * PI * fc))) * std::exp(-1 * std::pow(2.0 * PI * fc * t / 3.0 - 3, 2)));
			t += 0.5 * dT;
			s->sig[1][n] = (float)(std::cos(2.0 * PI * f0 * (t - 9.0 / (2.0 * PI * fc))) * std::exp(-1 * std::pow(2.0 * PI * fc * t / 3.0 - 3, 2)));
		}
		s->nyquist = calc_nyquist(f0 + fc, dT);
	} else {
		const double f0 = s->exc_f0;
		const unsigned len = (unsigned)std::round(2.0 / f0 / dT);
		s->sig_len = len;
		s->sig[0].assign(len, 0.f); s->sig[1].assign(len, 0.f);
		for (unsigned n = 1; n < len; ++n) {
			double t = n * dT;
			s->sig[0][n] = (float)std::sin(2.0 * PI * f0 * t);
			t += 0.5 * dT;
			s->sig[1][n] = (float)std::sin(2.0 * PI * f0 * t);
		}
		s->nyquist = calc_nyquist(f0, dT);
	}

	// ---- compressed operator, one representative plane per class
	const size_t np = (size_t)Nx * Ny;
	std::vector<std::vector<uint32_t>> plane_index(rep.size());
	const int nthreads = std::max(1, omp_get_max_threads());
	for (size_t r = 0; r < rep.size(); ++r) {
		const unsigned k = rep[r];
		PlaneEC E;
		s->ec_plane(k, E);
		std::vector<oems_coeff_entry> ent(np);
#pragma omp parallel for schedule(static) num_threads(nthreads)
		for (long long j = 0; j < (long long)Ny; ++j)
			for (unsigned i = 0; i < Nx; ++i) {
				const unsigned pos[3] = {i, (unsigned)j, k};
				s->cell_entry(pos, E, any_metal, any_exc, ent[(size_t)j * Nx + i]);
			}
		// de-duplicate: consecutive cells are mostly equal, so test the previous one first
		plane_index[r].resize(np);
		uint32_t last_id = 0;
		const oems_coeff_entry* last = nullptr;
		for (size_t p = 0; p < np; ++p) {
			if (last && memcmp(last, &ent[p], sizeof(oems_coeff_entry)) == 0) { plane_index[r][p] = last_id; continue; }
			last_id = s->table.insert(ent[p]);
			last = &ent[p];
			plane_index[r][p] = last_id;
		}
	}
	const unsigned U = (unsigned)s->table.items.size();
	s->index_bytes = (U + 1 <= 65536) ? 2 : 4;
	s->plane_of_z.assign(plane_id.begin(), plane_id.end());
	if (s->index_bytes == 2) {
		s->uplanes16.resize(np * rep.size());
		for (size_t r = 0; r < rep.size(); ++r) {
			uint16_t* dst = s->uplanes16.data() + r * np;
			for (size_t p = 0; p < np; ++p) dst[p] = (uint16_t)plane_index[r][p];
		}
	} else {
		s->uplanes32.resize(np * rep.size());
		for (size_t r = 0; r < rep.size(); ++r) memcpy(s->uplanes32.data() + r * np, plane_index[r].data(), np * 4);
	}

	// ---- excitation lists: operator_ext_excitation.cpp:143-232, loop order z, y, x
	if (any_exc) {
		// only the index range covered by excitation boxes needs a visit
		unsigned lo[3] = {UINT_MAX, UINT_MAX, UINT_MAX}, hi[3] = {0, 0, 0};
		for (const Prop& p : s->props) {
			if (p.type != P_EXCITATION) continue;
			for (int a = 0; a < 3; ++a) {
				unsigned l = 0, h = s->N[a] - 1;
				while (l + 1 < s->N[a] && s->disc_line(a, l + 1, false) < p.start[a]) ++l;
				while (h > 0 && s->disc_line(a, h - 1, false) > p.stop[a]) --h;
				lo[a] = std::min(lo[a], l); hi[a] = std::max(hi[a], h);
			}
		}
		unsigned pos[3];
		double c[3];
		for (pos[2] = lo[2]; pos[2] <= hi[2]; ++pos[2])
			for (pos[1] = lo[1]; pos[1] <= hi[1]; ++pos[1])
				for (pos[0] = lo[0]; pos[0] <= hi[0]; ++pos[0]) {
					for (int n = 0; n < 3; ++n) {
						if (!s->yee_coords(n, pos, c, false)) continue;
						const Prop* e = s->prop_at(c, MASK_EXC);
						if (!e) continue;
						if (e->exc_type == 0 || e->exc_type == 1) {
							const double amp = e->exc_vec[n] * s->edge_length(n, pos, false);
							if (amp != 0) {
								for (int a = 0; a < 3; ++a) s->exc_idx[0][a].push_back(pos[a]);
								s->exc_dir[0].push_back(n); s->exc_amp[0].push_back((float)amp);
								s->exc_delay[0].push_back((unsigned)(e->delay / dT));
							}
						}
					}
					for (int n = 0; n < 3; ++n) {
						if (pos[0] >= Nx - 1 || pos[1] >= Ny - 1 || pos[2] >= Nz - 1) continue;
						if (!s->yee_coords(n, pos, c, true)) continue;
						const Prop* e = s->prop_at(c, MASK_EXC);
						if (!e) continue;
						if (e->exc_type == 2 || e->exc_type == 3) {
							const double amp = e->exc_vec[n] * s->edge_length(n, pos, true);
							if (amp != 0) {
								for (int a = 0; a < 3; ++a) s->exc_idx[1][a].push_back(pos[a]);
								s->exc_dir[1].push_back(n); s->exc_amp[1].push_back((float)amp);
								s->exc_delay[1].push_back((unsigned)(e->delay / dT));
							}
						}
					}
				}
	}

	// ---- Mur planes: operator_ext_mur_abc.cpp:80-186, engine_ext_mur_abc.cpp:44-60
	for (int f = 0; f < 6; ++f) {
		if (s->bc[f] != 2) continue;
		MurPlaneH M;
		const int ny = f / 2, nyP = (ny + 1) % 3, nyPP = (ny + 2) % 3;
		const bool top = f % 2;
		M.ny = ny;
		M.line = top ? s->N[ny] - 1 : 0;
		M.shift = top ? s->N[ny] - 2 : 1;
		M.n[0] = s->N[nyP]; M.n[1] = s->N[nyPP];
		M.cP.resize((size_t)M.n[0] * M.n[1]); M.cPP.resize(M.cP.size());
		unsigned pos[3] = {0, 0, 0};
		pos[ny] = M.line;
		const double delta = std::fabs(s->edge_length(ny, pos, false));
		double coord[3];
		coord[ny] = M.line == 0 ? s->disc_line(ny, pos[ny], false) + delta / 2 / s->gd : s->disc_line(ny, pos[ny], false) - delta / 2 / s->gd;
		for (pos[nyP] = 0; pos[nyP] < M.n[0]; ++pos[nyP]) {
			coord[nyP] = s->disc_line(nyP, pos[nyP], false);
			for (pos[nyPP] = 0; pos[nyPP] < M.n[1]; ++pos[nyPP]) {
				coord[nyPP] = s->disc_line(nyPP, pos[nyPP], false);
				const Prop* p = s->prop_at(coord, MASK_MAT);
				double c0t;
				const size_t o = (size_t)pos[nyP] * M.n[1] + pos[nyPP];
				if (p) c0t = C0 * dT / std::sqrt(p->epsR * p->mueR);
				else c0t = C0 / std::sqrt(s->bg[0] * s->bg[1]) * dT;
				M.cP[o] = (float)((c0t - delta) / (c0t + delta));
				M.cPP[o] = M.cP[o];
			}
		}
		int maxDelay = -1;
		for (size_t n = 0; n < s->exc_dir[0].size(); ++n)
			if (((int)s->exc_dir[0][n] == nyP || (int)s->exc_dir[0][n] == nyPP) && s->exc_idx[0][ny][n] == M.line)
				maxDelay = std::max(maxDelay, (int)s->exc_delay[0][n]);
		M.start_ts = maxDelay >= 0 ? (unsigned)maxDelay + s->sig_len + 10 : 0;
		s->mur.push_back(std::move(M));
	}

	// ---- Lorentz/Drude lists: operator_ext_lorentzmaterial.cpp:120-445 over the planes that
	// touch a Lorentz box
	s->lor_order = 0;
	if (any_lor) {
		for (const Prop& p : s->props)
			if (p.type == P_LORENTZ) s->lor_order = std::max(s->lor_order, p.order);
		s->lor.resize(s->lor_order);
		unsigned lo[3] = {UINT_MAX, UINT_MAX, UINT_MAX}, hi[3] = {0, 0, 0};
		for (const Prop& p : s->props) {
			if (p.type != P_LORENTZ) continue;
			for (int a = 0; a < 3; ++a) {
				unsigned l = 0, h = s->N[a] - 1;
				while (l + 1 < s->N[a] && s->disc_line(a, l + 1, false) < p.start[a]) ++l;
				while (h > 0 && s->disc_line(a, h - 1, false) > p.stop[a]) --h;
				lo[a] = std::min(lo[a], l); hi[a] = std::max(hi[a], h);
			}
		}
		const oems_coeff_entry* tab = s->table.items.data();
		for (int order = 0; order < s->lor_order; ++order) {
			LorOrderH& Lo = s->lor[order];
			// flags first (they are global in the reference and only switch on)
			// per-plane lists, concatenated in x-major order afterwards
			struct Item { unsigned pos[3]; float c[6][3]; };
			std::vector<std::vector<Item>> per_x(hi[0] - lo[0] + 1);
			bool volt_on = false, curr_on = false, volt_lor = false, curr_lor = false;
			// EC values are needed: recompute the EC planes in the z range
			std::vector<PlaneEC> ecs(hi[2] - lo[2] + 1);
			for (unsigned k = lo[2]; k <= hi[2]; ++k) s->ec_plane(k, ecs[k - lo[2]]);
			for (unsigned i = lo[0]; i <= hi[0]; ++i)
				for (unsigned j = lo[1]; j <= hi[1]; ++j)
					for (unsigned k = lo[2]; k <= hi[2]; ++k) {
						const unsigned pos[3] = {i, j, k};
						const size_t p2 = (size_t)j * Nx + i;
						const PlaneEC& E = ecs[k - lo[2]];
						const uint32_t id = s->cell_id(k, p2);
						const oems_coeff_entry& ce = tab[id];
						bool on = false;
						double L_D[3], R_D[3], C_L[3], C_D[3], G_D[3], L_L[3], coord[3];
						for (int n = 0; n < 3; ++n) {
							L_D[n] = R_D[n] = C_L[n] = 0;
							if (!s->yee_coords(n, pos, coord, false)) continue;
							if (ce.vi[n] == 0) continue;
							const Prop* q = s->prop_at(coord, MASK_MAT_METAL);
							if (!q || q->type != P_LORENTZ) continue;
							const double wp = (order < q->order ? q->eps_fp[order] : 0) * 2 * PI;
							if (wp > 0 && E.C[n][p2] > 0) { on = true; volt_on = true; L_D[n] = 1 / (wp * wp * E.C[n][p2]); }
							const double tr = order < q->order ? q->eps_tau[order] : 0;
							if (tr > 0 && volt_on) R_D[n] = L_D[n] / tr;
							const double wl = (order < q->order ? q->eps_flor[order] : 0) * 2 * PI;
							if (wl > 0 && L_D[n] > 0) { volt_lor = true; C_L[n] = 1 / (wl * wl * L_D[n]); }
						}
						for (int n = 0; n < 3; ++n) {
							C_D[n] = G_D[n] = L_L[n] = 0;
							if (!s->yee_coords(n, pos, coord, true)) continue;
							if (ce.iv[n] == 0) continue;
							const Prop* q = s->prop_at(coord, MASK_MAT_METAL);
							if (!q || q->type != P_LORENTZ) continue;
							const double wp = (order < q->order ? q->mue_fp[order] : 0) * 2 * PI;
							if (wp > 0 && E.L[n][p2] > 0) { on = true; curr_on = true; C_D[n] = 1 / (wp * wp * E.L[n][p2]); }
							const double tr = order < q->order ? q->mue_tau[order] : 0;
							if (tr > 0 && curr_on) G_D[n] = C_D[n] / tr;
							const double wl = (order < q->order ? q->mue_flor[order] : 0) * 2 * PI;
							if (wl > 0 && C_D[n] > 0) { curr_lor = true; L_L[n] = 1 / (wl * wl * C_D[n]); }
						}
						if (!on) continue;
						Item it;
						for (int a = 0; a < 3; ++a) it.pos[a] = pos[a];
						for (int n = 0; n < 3; ++n) {
							const double VI = ce.vi[n], IV = ce.iv[n];
							if (L_D[n] > 0) {
								it.c[0][n] = (float)((2.0 * L_D[n] - dT * R_D[n]) / (2.0 * L_D[n] + dT * R_D[n]));
								it.c[1][n] = (float)(dT / (L_D[n] + dT * R_D[n] / 2.0) * VI);
							} else if (R_D[n] > 0 && C_L[n] > 0) {
								it.c[0][n] = (float)((2.0 * dT - R_D[n] * C_L[n]) / (C_L[n] * R_D[n]));
								it.c[1][n] = (float)(2.0 / R_D[n] * VI);
							} else { it.c[0][n] = 1; it.c[1][n] = 0; }
							if (C_D[n] > 0) {
								it.c[3][n] = (float)((2.0 * C_D[n] - dT * G_D[n]) / (2.0 * C_D[n] + dT * G_D[n]));
								it.c[4][n] = (float)(dT / (C_D[n] + dT * G_D[n] / 2.0) * IV);
							} else { it.c[3][n] = 1; it.c[4][n] = 0; }
							it.c[2][n] = C_L[n] > 0 ? (float)(dT / C_L[n] / VI) : 0;
							it.c[5][n] = L_L[n] > 0 ? (float)(dT / L_L[n] / IV) : 0;
						}
						per_x[i - lo[0]].push_back(it);
					}
			Lo.volt_on = volt_on; Lo.curr_on = curr_on; Lo.volt_lor_on = volt_lor; Lo.curr_lor_on = curr_lor;
			for (auto& v : per_x)
				for (const Item& it : v) {
					for (int a = 0; a < 3; ++a) Lo.pos[a].push_back(it.pos[a]);
					for (int w = 0; w < 6; ++w)
						for (int n = 0; n < 3; ++n) Lo.c[w][n].push_back(it.c[w][n]);
				}
		}
	}
	s->built = true;
	return 0;
}

double oems_synth_dT(const oems_synth* s) { return s->dT; }
unsigned oems_synth_nyquist(const oems_synth* s) { return s->nyquist; }
unsigned oems_synth_n_unique(const oems_synth* s) { return (unsigned)s->table.items.size(); }
int oems_synth_index_bytes(const oems_synth* s) { return s->index_bytes; }
const oems_coeff_entry* oems_synth_table(const oems_synth* s) { return s->table.items.data(); }
const void* oems_synth_index(const oems_synth* cs)
{
	oems_synth* s = const_cast<oems_synth*>(cs);
	s->ensure_full_index();
	return s->index_bytes == 2 ? (const void*)s->idx16.data() : (const void*)s->idx32.data();
}
unsigned oems_synth_unique_planes(const oems_synth* s) { return s->unique_planes; }
const unsigned* oems_synth_plane_of_z(const oems_synth* s) { return s->plane_of_z.data(); }
const void* oems_synth_plane_data(const oems_synth* s) { return s->index_bytes == 2 ? (const void*)s->uplanes16.data() : (const void*)s->uplanes32.data(); }
unsigned oems_synth_signal_length(const oems_synth* s) { return s->sig_len; }
const float* oems_synth_signal(const oems_synth* s, int is_curr) { return s->sig[is_curr ? 1 : 0].data(); }
unsigned oems_synth_exc_count(const oems_synth* s, int w) { return (unsigned)s->exc_dir[w ? 1 : 0].size(); }
void oems_synth_exc_get(const oems_synth* s, int w, unsigned* idx3, unsigned* dir, float* amp, unsigned* delay)
{
	w = w ? 1 : 0;
	const size_t cnt = s->exc_dir[w].size();
	for (int a = 0; a < 3; ++a) memcpy(idx3 + a * cnt, s->exc_idx[w][a].data(), cnt * sizeof(unsigned));
	memcpy(dir, s->exc_dir[w].data(), cnt * sizeof(unsigned));
	memcpy(amp, s->exc_amp[w].data(), cnt * sizeof(float));
	memcpy(delay, s->exc_delay[w].data(), cnt * sizeof(unsigned));
}
int oems_synth_mur_count(const oems_synth* s) { return (int)s->mur.size(); }
const float* oems_synth_mur_coeff(const oems_synth* s, int m, int which, int* ny, unsigned* line, unsigned* shift, unsigned nl[2], unsigned* start_ts)
{
	const MurPlaneH& M = s->mur[m];
	if (ny) *ny = M.ny;
	if (line) *line = M.line;
	if (shift) *shift = M.shift;
	if (nl) { nl[0] = M.n[0]; nl[1] = M.n[1]; }
	if (start_ts) *start_ts = M.start_ts;
	return which ? M.cPP.data() : M.cP.data();
}
int oems_synth_upml_count(const oems_synth* s) { return (int)s->upml.size(); }
void oems_synth_upml_box(const oems_synth* s, int b, unsigned start[3], unsigned nl[3])
{
	for (int a = 0; a < 3; ++a) { start[a] = s->upml[b].start[a]; nl[a] = s->upml[b].n[a]; }
}
int oems_synth_lorentz_order(const oems_synth* s) { return s->lor_order; }
unsigned oems_synth_lorentz_count(const oems_synth* s, int o) { return (unsigned)s->lor[o].pos[0].size(); }

// page-locks the per-cell index so that oems_cuda_set_operator_compressed can DMA it straight
// from the builder's buffer (PCIe rate) instead of staging it through a pinned bounce buffer
int oems_synth_pin(oems_synth* s)
{
	if (!s || !s->built) return 1;
	if (s->pinned) return 0;
	s->ensure_full_index();
	const size_t bytes = (size_t)s->index_bytes * (s->index_bytes == 2 ? s->idx16.size() : s->idx32.size());
	if (cudaHostRegister(const_cast<void*>(oems_synth_index(s)), bytes, cudaHostRegisterPortable) != cudaSuccess) {
		cudaGetLastError();
		return 2;
	}
	s->pinned = true;
	return 0;
}

int oems_synth_upload(const oems_synth* s, oems_cuda_engine* eng)
{
	if (!s || !s->built || !eng) return 1;
	// unique xy planes + a plane id per z: for meshes that repeat along z this is a small fraction of the
	// full per-cell index (C5 1024^3: 20 planes of 1024), the engine expands it on the device
	int rc = oems_cuda_set_operator_planes(eng, oems_synth_n_unique(s), oems_synth_table(s), s->unique_planes,
	                                       s->index_bytes == 2 ? (const void*)s->uplanes16.data() : (const void*)s->uplanes32.data(),
	                                       s->plane_of_z.data(), s->index_bytes);
	if (rc) return rc;
	rc = oems_cuda_set_signal(eng, s->sig[0].data(), s->sig[1].data(), s->sig_len,
	                          s->exc_period > 0 ? (unsigned)(int)(s->exc_period / s->dT) : 0);
	if (rc) return rc;
	for (int w = 0; w < 2; ++w) {
		const unsigned cnt = (unsigned)s->exc_dir[w].size();
		if (!cnt) continue;
		std::vector<unsigned> idx3((size_t)3 * cnt);
		for (int a = 0; a < 3; ++a) memcpy(idx3.data() + (size_t)a * cnt, s->exc_idx[w][a].data(), cnt * sizeof(unsigned));
		rc = oems_cuda_add_excitation(eng, w, cnt, idx3.data(), s->exc_dir[w].data(), s->exc_amp[w].data(), s->exc_delay[w].data());
		if (rc) return rc;
	}
	for (const MurPlaneH& M : s->mur) {
		rc = oems_cuda_add_mur(eng, M.ny, M.line, M.shift, M.n, M.cP.data(), M.cPP.data(), M.start_ts);
		if (rc) return rc;
	}
	for (const UpmlBoxH& B : s->upml) {
		rc = oems_cuda_add_upml(eng, B.start, B.n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
		if (rc) return rc;
	}
	for (const LorOrderH& L : s->lor) {
		const unsigned cnt = (unsigned)L.pos[0].size();
		std::vector<unsigned> pos3((size_t)3 * cnt);
		std::vector<float> c[6];
		for (int a = 0; a < 3; ++a) memcpy(pos3.data() + (size_t)a * cnt, L.pos[a].data(), cnt * sizeof(unsigned));
		for (int w = 0; w < 6; ++w) {
			c[w].resize((size_t)3 * cnt);
			for (int n = 0; n < 3; ++n) memcpy(c[w].data() + (size_t)n * cnt, L.c[w][n].data(), cnt * sizeof(float));
		}
		rc = oems_cuda_add_lorentz(eng, cnt, pos3.data(), L.volt_on ? c[0].data() : nullptr, L.volt_on ? c[1].data() : nullptr,
		                           L.volt_lor_on ? c[2].data() : nullptr, L.curr_on ? c[3].data() : nullptr,
		                           L.curr_on ? c[4].data() : nullptr, L.curr_lor_on ? c[5].data() : nullptr);
		if (rc) return rc;
	}
	return oems_cuda_finalize(eng);
}

} // extern "C"
