/*
 * fdtd_oracle.c -- CPU restatement of the openEMS FDTD hot path.  TEST INFRASTRUCTURE ONLY
 * (see fdtd_oracle.h for the rules and the pinning statement).
 *
 * Build: gcc -O2 -std=gnu11 -ffp-contract=off -msse2 -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off: the reference is built for baseline x86-64 (CMakeLists.txt:4-8, no -march)
 * so every fp32 multiply and add is rounded separately; FTZ/DAZ is set while stepping
 * (tools/denormal.h:19-30, FDTD/engine_sse.cpp:43).
 */
#include "fdtd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <xmmintrin.h>

#include "fdtd_oracle_priv.h"

/* ---------------------------------------------------------------- helpers */
#define IDX orc_idx
static size_t ncell(const orc_sim* s) { return (size_t)s->N[0] * s->N[1] * s->N[2]; }
static void* xcalloc(size_t n, size_t sz)
{
	void* p = calloc(n ? n : 1, sz);
	if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
	return p;
}

/* Operator::GetDiscLine operator.cpp:143-157 */
double orc_disc_line(const orc_sim* s, int n, unsigned pos, int dual)
{
	if (n < 0 || n > 2) return 0.0;
	if (pos >= s->N[n]) return 0.0;
	const double* L = s->lines[n];
	if (!dual) return L[pos];
	if (pos < s->N[n] - 1) return 0.5 * (L[pos] + L[pos + 1]);
	return L[pos] + 0.5 * (L[pos] - L[pos - 1]);
}
/* Operator::GetDiscDelta operator.cpp:159-180 */
static double disc_delta(const orc_sim* s, int n, unsigned pos, int dual)
{
	if (n < 0 || n > 2) return 0.0;
	if (pos >= s->N[n]) return 0.0;
	if (!dual) {
		if (pos < s->N[n] - 1) return orc_disc_line(s, n, pos + 1, 0) - orc_disc_line(s, n, pos, 0);
		return orc_disc_line(s, n, pos, 0) - orc_disc_line(s, n, pos - 1, 0);
	}
	if (pos > 0) return orc_disc_line(s, n, pos, 1) - orc_disc_line(s, n, pos - 1, 1);
	return orc_disc_line(s, n, 1, 0) - orc_disc_line(s, n, 0, 0);
}
/* Operator::GetEdgeLength operator.cpp:208-211 */
double orc_edge_length(const orc_sim* s, int n, const unsigned pos[3], int dual)
{
	return disc_delta(s, n, pos[n], dual) * s->grid_delta;
}
/* Operator::GetNodeWidth operator.h:174 -> GetEdgeLength(ny,pos,!dualMesh) */
static double node_width(const orc_sim* s, int ny, const unsigned pos[3], int dual)
{
	return orc_edge_length(s, ny, pos, !dual);
}
/* Operator::GetNodeArea operator.cpp:235-240 ; GetEdgeArea operator.h:194 */
static double node_area(const orc_sim* s, int ny, const unsigned pos[3], int dual)
{
	return node_width(s, (ny + 1) % 3, pos, dual) * node_width(s, (ny + 2) % 3, pos, dual);
}
static double edge_area(const orc_sim* s, int ny, const unsigned pos[3], int dual)
{
	return node_area(s, ny, pos, dual);
}
/* Operator::GetYeeCoords operator.cpp:182-203 */
static int yee_coords(const orc_sim* s, int ny, const unsigned pos[3], double* c, int dual)
{
	for (int n = 0; n < 3; ++n) c[n] = orc_disc_line(s, n, pos[n], dual);
	c[ny] = orc_disc_line(s, ny, pos[ny], !dual);
	if (!dual) {
		if (pos[ny] >= s->N[ny] - 1) return 0;
	} else {
		int nP = (ny + 1) % 3, nPP = (ny + 2) % 3;
		if (pos[nP] >= s->N[nP] - 1 || pos[nPP] >= s->N[nPP] - 1) return 0;
	}
	return 1;
}

/* stand-in for ContinuousStructure::GetPropertyByCoordPriority (CSXCAD, not vendored):
   highest priority box of an accepted type containing the point, ties -> added later */
static const prop_t* prop_at(const orc_sim* s, const double c[3], unsigned mask)
{
	const prop_t* best = NULL;
	for (int p = 0; p < s->nprops; ++p) {
		const prop_t* q = &s->props[p];
		if (!(mask & (1u << q->type))) continue;
		if (c[0] < q->start[0] || c[0] > q->stop[0] || c[1] < q->start[1] || c[1] > q->stop[1] ||
		    c[2] < q->start[2] || c[2] > q->stop[2])
			continue;
		if (!best || q->prio >= best->prio) best = q;
	}
	return best;
}
#define MASK_MAT ((1u << P_MATERIAL) | (1u << P_LORENTZ))
#define MASK_MAT_METAL (MASK_MAT | (1u << P_METAL))
#define MASK_EXC (1u << P_EXCITATION)

/* Operator::GetMaterial operator.cpp:1289-1345 */
static double get_material(const orc_sim* s, const double c[3], int type)
{
	const prop_t* p = prop_at(s, c, MASK_MAT);
	if (p) {
		switch (type) {
		case 0: return p->epsR;
		case 1: return p->kappa;
		case 2: return p->mueR;
		case 3: return p->sigma;
		}
	}
	return s->bg[type == 0 ? 0 : type == 1 ? 2 : type == 2 ? 1 : 3];
}
/* bg layout: [0]=epsR [1]=mueR [2]=kappa [3]=sigma */

/* Operator::GetCellCenterMaterialAvgCoord operator.cpp:1277-1287 */
static int cell_center(const orc_sim* s, const int pos[3], double c[3])
{
	for (int n = 0; n < 3; ++n)
		if (pos[n] < 0 || pos[n] >= (int)s->N[n]) return 0;
	for (int n = 0; n < 3; ++n) c[n] = orc_disc_line(s, n, (unsigned)pos[n], 1);
	return 1;
}

/* Operator::AverageMatCellCenter operator.cpp:1347-1444 (CellConstantMaterial averaging; for
   boxes aligned to mesh lines it coincides with the default quarter-cell averaging) */
static void eff_mat(const orc_sim* s, int ny, const unsigned pos[3], double E[4])
{
	int n = ny, nP = (n + 1) % 3, nPP = (n + 2) % 3;
	int lp[3] = {(int)pos[0], (int)pos[1], (int)pos[2]};
	double c[3], A, area = 0;
	E[0] = E[1] = E[2] = E[3] = 0;
	unsigned up[3];
#define ACC_EPS()                                                                         \
	if (cell_center(s, lp, c)) {                                                          \
		up[0] = lp[0]; up[1] = lp[1]; up[2] = lp[2];                                      \
		A = node_area(s, ny, up, 1);                                                      \
		E[0] += get_material(s, c, 0) * A;                                                \
		E[1] += get_material(s, c, 1) * A;                                                \
		area += A;                                                                        \
	}
	ACC_EPS();
	--lp[nP];
	ACC_EPS();
	++lp[nP]; --lp[nPP];
	ACC_EPS();
	--lp[nP];
	ACC_EPS();
#undef ACC_EPS
	E[0] *= EPS0 / area;
	E[1] /= area;

	lp[0] = pos[0]; lp[1] = pos[1]; lp[2] = pos[2];
	double length = 0, d, sig;
	--lp[n];
	if (cell_center(s, lp, c)) {
		up[0] = lp[0]; up[1] = lp[1]; up[2] = lp[2];
		d = node_width(s, n, up, 1);
		E[2] += d / get_material(s, c, 2);
		sig = get_material(s, c, 3);
		if (sig) E[3] += d / sig; else E[3] = 0;
		length += d;
	}
	++lp[n];
	if (cell_center(s, lp, c)) {
		up[0] = lp[0]; up[1] = lp[1]; up[2] = lp[2];
		d = node_width(s, n, up, 1);
		E[2] += d / get_material(s, c, 2);
		sig = get_material(s, c, 3);
		if (sig) E[3] += d / sig; else E[3] = 0;
		length += d;
	}
	E[2] = length * MUE0 / E[2];
	if (E[3]) E[3] = length / E[3];
}

/* Operator::Calc_ECPos operator.cpp:1189-1256 */
static void calc_ec_pos(const orc_sim* s, int ny, const unsigned pos[3], double EC[4])
{
	double E[4];
	eff_mat(s, ny, pos, E);
	double delta = orc_edge_length(s, ny, pos, 0), area = edge_area(s, ny, pos, 0);
	if (delta) { EC[0] = E[0] * area / delta; EC[1] = E[1] * area / delta; }
	else { EC[0] = 0; EC[1] = 0; }
	delta = orc_edge_length(s, ny, pos, 1);
	area = edge_area(s, ny, pos, 1);
	if (delta) { EC[2] = E[2] * area / delta; EC[3] = E[3] * area / delta; }
	else { EC[2] = 0; EC[3] = 0; }
}

/* ---------------------------------------------------------------- public setup */
orc_sim* orc_create(const unsigned nl[3], const double* x, const double* y, const double* z,
                    double grid_delta)
{
	orc_sim* s = xcalloc(1, sizeof(*s));
	const double* src[3] = {x, y, z};
	for (int n = 0; n < 3; ++n) {
		if (nl[n] < 3) { free(s); return NULL; } /* operator.cpp:798-803 */
		s->N[n] = nl[n];
		s->lines[n] = xcalloc(nl[n], sizeof(double));
		memcpy(s->lines[n], src[n], nl[n] * sizeof(double));
	}
	s->grid_delta = grid_delta;
	s->bg[0] = 1; s->bg[1] = 1; s->bg[2] = 0; s->bg[3] = 0;
	for (int n = 0; n < 6; ++n) s->pml_size[n] = 8; /* openems.cpp:129 */
	s->ts_factor = 1.0;
	s->exc_kind = -1;
	return s;
}

static void free_upml(upml_t* u)
{
	for (int c = 0; c < 6; ++c) free(u->c[c]);
	free(u->volt_flux); free(u->curr_flux);
}

void orc_destroy(orc_sim* s)
{
	if (!s) return;
	for (int n = 0; n < 3; ++n) free(s->lines[n]);
	free(s->props); free(s->lumped);
	free(s->EC_C); free(s->EC_G); free(s->EC_L); free(s->EC_R);
	free(s->vv); free(s->vi); free(s->ii); free(s->iv);
	free(s->sig_v); free(s->sig_i);
	for (int n = 0; n < 3; ++n) { free(s->vidx[n]); free(s->cidx[n]); }
	free(s->vdir); free(s->vdelay); free(s->vamp);
	free(s->cdir); free(s->cdelay); free(s->camp);
	for (int b = 0; b < s->nupml; ++b) free_upml(&s->upml[b]);
	for (int m = 0; m < s->nmur; ++m) {
		free(s->mur[m].cP); free(s->mur[m].cPP); free(s->mur[m].vP); free(s->mur[m].vPP);
	}
	for (int o = 0; o < s->lor_order; ++o) {
		lor_order_t* L = &s->lor[o];
		for (int n = 0; n < 3; ++n) {
			free(L->pos[n]); free(L->v_int[n]); free(L->v_ext[n]); free(L->v_lor[n]);
			free(L->i_int[n]); free(L->i_ext[n]); free(L->i_lor[n]);
			free(L->volt_ADE[n]); free(L->curr_ADE[n]); free(L->volt_Lor_ADE[n]); free(L->curr_Lor_ADE[n]);
		}
	}
	for (int r = 0; r < s->nrlc; ++r) {
		rlc_t* R = &s->rlc[r];
		free(R->dir);
		for (int n = 0; n < 3; ++n) { free(R->pos[n]); free(R->Vdn[n]); free(R->Jn[n]); }
		free(R->ilv); free(R->i2v); free(R->vvd); free(R->vv2); free(R->vj1); free(R->vj2);
		free(R->ib0); free(R->b1); free(R->b2); free(R->Il);
	}
	free(s->rlc);
	if (s->ss) { free(s->ss->dir); for (int n = 0; n < 3; ++n) free(s->ss->pos[n]); free(s->ss->rec); free(s->ss); }
	free(s->volt); free(s->curr);
	free(s);
}

void orc_set_bc(orc_sim* s, const int bc[6], const unsigned pml_size[6])
{
	for (int n = 0; n < 6; ++n) {
		s->bc[n] = bc[n];
		if (pml_size) s->pml_size[n] = pml_size[n];
	}
}
void orc_set_background(orc_sim* s, double epsR, double mueR, double kappa, double sigma)
{
	s->bg[0] = epsR; s->bg[1] = mueR; s->bg[2] = kappa; s->bg[3] = sigma;
}
void orc_set_mur_phase_velocity(orc_sim* s, double v) { s->mur_vphase = v; }
void orc_set_timestep(orc_sim* s, double forced_dT, double factor)
{
	s->forced_dT = forced_dT;
	s->ts_factor = factor > 0 ? factor : 1.0;
}

static prop_t* new_prop(orc_sim* s, int type, int prio, const double a[3], const double b[3])
{
	s->props = realloc(s->props, (size_t)(s->nprops + 1) * sizeof(prop_t));
	prop_t* p = &s->props[s->nprops++];
	memset(p, 0, sizeof(*p));
	p->type = type; p->prio = prio;
	for (int n = 0; n < 3; ++n) {
		p->start[n] = a[n] < b[n] ? a[n] : b[n];
		p->stop[n] = a[n] < b[n] ? b[n] : a[n];
	}
	p->epsR = 1; p->mueR = 1;
	return p;
}
int orc_add_material(orc_sim* s, int prio, const double a[3], const double b[3], double epsR,
                     double mueR, double kappa, double sigma)
{
	prop_t* p = new_prop(s, P_MATERIAL, prio, a, b);
	p->epsR = epsR; p->mueR = mueR; p->kappa = kappa; p->sigma = sigma;
	return s->nprops - 1;
}
int orc_add_metal(orc_sim* s, int prio, const double a[3], const double b[3])
{
	new_prop(s, P_METAL, prio, a, b);
	return s->nprops - 1;
}
int orc_add_lorentz(orc_sim* s, int prio, const double a[3], const double b[3], double epsR,
                    double mueR, double kappa, double sigma, int order, const double* eps_fp,
                    const double* eps_tau, const double* eps_flor, const double* mue_fp,
                    const double* mue_tau, const double* mue_flor)
{
	if (order > MAX_ORDER) return -1;
	prop_t* p = new_prop(s, P_LORENTZ, prio, a, b);
	p->epsR = epsR; p->mueR = mueR; p->kappa = kappa; p->sigma = sigma;
	p->order = order;
	for (int o = 0; o < order; ++o) {
		p->eps_fp[o] = eps_fp ? eps_fp[o] : 0; p->eps_tau[o] = eps_tau ? eps_tau[o] : 0;
		p->eps_flor[o] = eps_flor ? eps_flor[o] : 0;
		p->mue_fp[o] = mue_fp ? mue_fp[o] : 0; p->mue_tau[o] = mue_tau ? mue_tau[o] : 0;
		p->mue_flor[o] = mue_flor ? mue_flor[o] : 0;
	}
	return s->nprops - 1;
}
int orc_add_excitation(orc_sim* s, int prio, const double a[3], const double b[3], int exc_type,
                       const double vec[3], double delay_s)
{
	prop_t* p = new_prop(s, P_EXCITATION, prio, a, b);
	p->exc_type = exc_type;
	for (int n = 0; n < 3; ++n) p->exc_vec[n] = vec[n];
	p->delay = delay_s;
	return s->nprops - 1;
}
int orc_add_lumped_rc(orc_sim* s, const double a[3], const double b[3], int dir, double R,
                      double C, int caps)
{
	s->lumped = realloc(s->lumped, (size_t)(s->nlumped + 1) * sizeof(lumped_t));
	lumped_t* l = &s->lumped[s->nlumped++];
	for (int n = 0; n < 3; ++n) {
		l->start[n] = a[n] < b[n] ? a[n] : b[n];
		l->stop[n] = a[n] < b[n] ? b[n] : a[n];
	}
	l->dir = dir; l->R = R; l->C = C; l->caps = caps;
	return s->nlumped - 1;
}

static float* dupf(const float* src, unsigned n)
{
	float* d = xcalloc(n, sizeof(float));
	if (src) memcpy(d, src, n * sizeof(float));
	return d;
}
int orc_add_rlc_raw(orc_sim* s, unsigned count, const int* dir, const unsigned* pos,
                    const float* ilv, const float* i2v, const float* vvd, const float* vv2,
                    const float* vj1, const float* vj2, const float* ib0, const float* b1,
                    const float* b2)
{
	s->rlc = realloc(s->rlc, (size_t)(s->nrlc + 1) * sizeof(rlc_t));
	rlc_t* R = &s->rlc[s->nrlc++];
	memset(R, 0, sizeof(*R));
	R->count = count;
	R->dir = xcalloc(count, sizeof(int));
	memcpy(R->dir, dir, count * sizeof(int));
	for (int n = 0; n < 3; ++n) {
		R->pos[n] = xcalloc(count, sizeof(unsigned));
		memcpy(R->pos[n], pos + (size_t)n * count, count * sizeof(unsigned));
		R->Vdn[n] = xcalloc(count, sizeof(float));
		R->Jn[n] = xcalloc(count, sizeof(float));
	}
	R->Il = xcalloc(count, sizeof(float));
	R->ilv = dupf(ilv, count); R->i2v = dupf(i2v, count); R->vvd = dupf(vvd, count);
	R->vv2 = dupf(vv2, count); R->vj1 = dupf(vj1, count); R->vj2 = dupf(vj2, count);
	R->ib0 = dupf(ib0, count); R->b1 = dupf(b1, count); R->b2 = dupf(b2, count);
	return s->nrlc - 1;
}

void orc_set_excite_gauss(orc_sim* s, double f0, double fc)
{ /* Excitation::SetupGaussianPulse excitation.cpp:54-62 */
	s->exc_kind = 0; s->exc_f0 = f0; s->exc_fc = fc; s->exc_fmax = f0 + fc; s->exc_period = 0;
}
void orc_set_excite_sinus(orc_sim* s, double f0)
{ /* excitation.cpp:64-70 */
	s->exc_kind = 1; s->exc_f0 = f0; s->exc_fmax = f0; s->exc_period = 1 / f0;
}
void orc_set_excite_dirac(orc_sim* s, double fmax)
{ /* excitation.cpp:72-77 */
	s->exc_kind = 2; s->exc_fmax = fmax; s->exc_period = 0;
}
void orc_set_excite_step(orc_sim* s, double fmax)
{ /* excitation.cpp:79-84 */
	s->exc_kind = 3; s->exc_fmax = fmax; s->exc_period = 0;
}

/* ---------------------------------------------------------------- operator build */

/* tools/useful.cpp:30-36 */
static unsigned calc_nyquist(double fmax, double dT)
{
	if (fmax == 0) return UINT_MAX;
	if (dT == 0) return 1;
	double T0 = 1 / fmax;
	return (unsigned)floor(T0 / 2 / dT);
}

/* Excitation::buildExcitationSignal excitation.cpp:95-133 and the Calc* functions :150-276 */
static int build_signal(orc_sim* s, unsigned max_ts)
{
	double dT = s->dT;
	free(s->sig_v); free(s->sig_i); s->sig_v = s->sig_i = NULL;
	switch (s->exc_kind) {
	case 0: { /* CalcGaussianPulsExcitation excitation.cpp:150-176 */
		double f0 = s->exc_f0, fc = s->exc_fc;
		unsigned len = (unsigned)ceil(2.0 * 9.0 / (2.0 * ORC_PI * fc) / dT);
		if (len > max_ts) len = max_ts;
		s->sig_len = len;
		s->sig_v = xcalloc(len, sizeof(float));
		s->sig_i = xcalloc(len, sizeof(float));
		for (unsigned n = 0; n < len; ++n) {
			double t = n * dT;
			s->sig_v[n] = cos(2.0 * ORC_PI * f0 * (t - 9.0 / (2.0 * ORC_PI * fc))) *
			              exp(-1 * pow(2.0 * ORC_PI * fc * t / 3.0 - 3, 2));
			t += 0.5 * dT;
			s->sig_i[n] = cos(2.0 * ORC_PI * f0 * (t - 9.0 / (2.0 * ORC_PI * fc))) *
			              exp(-1 * pow(2.0 * ORC_PI * fc * t / 3.0 - 3, 2));
		}
		s->nyquist = calc_nyquist(f0 + fc, dT);
		break;
	}
	case 1: { /* CalcSinusExcitation excitation.cpp:254-276 */
		double f0 = s->exc_f0;
		unsigned len = (unsigned)round(2.0 / f0 / dT);
		s->sig_len = len;
		s->sig_v = xcalloc(len, sizeof(float));
		s->sig_i = xcalloc(len, sizeof(float));
		for (unsigned n = 1; n < len; ++n) {
			double t = n * dT;
			s->sig_v[n] = sin(2.0 * ORC_PI * f0 * t);
			t += 0.5 * dT;
			s->sig_i[n] = sin(2.0 * ORC_PI * f0 * t);
		}
		s->nyquist = calc_nyquist(f0, dT);
		break;
	}
	case 2: /* CalcDiracPulsExcitation excitation.cpp:178-197 */
		s->sig_len = 2;
		s->sig_v = xcalloc(2, sizeof(float)); s->sig_i = xcalloc(2, sizeof(float));
		s->sig_v[1] = 1.0f; s->sig_i[1] = 1.0f;
		s->nyquist = 1;
		break;
	case 3: /* CalcStepExcitation excitation.cpp:199-217 */
		s->sig_len = 2;
		s->sig_v = xcalloc(2, sizeof(float)); s->sig_i = xcalloc(2, sizeof(float));
		s->sig_v[0] = s->sig_v[1] = 1.0f; s->sig_i[0] = s->sig_i[1] = 1.0f;
		s->nyquist = 1;
		break;
	default:
		return -1;
	}
	return s->nyquist == 0 ? -1 : 0;
}

static double min4(const double* v)
{
	double m = v[0];
	for (int n = 1; n < 4; ++n) if (v[n] < m) m = v[n];
	return m;
}

/* AdrOp::GetPos with SetReflection2Cell, tools/AdrOp.cpp:183-222,429-433:
   an index that leaves [0,N) is mirrored about the cell (-1 -> 0, N -> N-1). */
static size_t refl_pos(const orc_sim* s, const unsigned pos[3], int di, int dj, int dk)
{
	int d[3] = {di, dj, dk};
	size_t p[3];
	for (int n = 0; n < 3; ++n) {
		int q = (int)pos[n] + d[n];
		if (q < 0) q = -q - 1;
		if (q > (int)s->N[n] - 1) q = 2 * ((int)s->N[n] - 1) - q + 1;
		p[n] = (size_t)q;
	}
	return (p[0] * s->N[1] + p[1]) * s->N[2] + p[2];
}

/* Operator::CalcTimestep_Var3 operator.cpp:1956-2030 (Rennings_2) */
static double calc_timestep(const orc_sim* s)
{
	double dT = 1e200;
	size_t nc = ncell(s);
	for (int n = 0; n < 3; ++n) {
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		const float *Cn = s->EC_C + n * nc, *CP = s->EC_C + nP * nc, *CPP = s->EC_C + nPP * nc;
		const float *LP = s->EC_L + nP * nc, *LPP = s->EC_L + nPP * nc;
		double dT_n = 1e200;
#pragma omp parallel for collapse(2) reduction(min : dT_n) schedule(static)
		for (unsigned k = 0; k < s->N[2]; ++k)
			for (unsigned j = 0; j < s->N[1]; ++j)
				for (unsigned i = 0; i < s->N[0]; ++i) {
					unsigned pos[3] = {i, j, k};
					int e[3][3] = {{0}};
					e[nP][nP] = 1; e[nPP][nPP] = 1; e[n][n] = 1;
#define SH(a, sa, b, sb) refl_pos(s, pos, e[a][0] * (sa) + e[b][0] * (sb), e[a][1] * (sa) + e[b][1] * (sb), e[a][2] * (sa) + e[b][2] * (sb))
					size_t ip = SH(n, 0, n, 0);
					/* EC_L and EC_C are FDTD_FLOAT arrays (operator.h:346-349): every product L*C, every 1/(L*C) and
					   every two-term sum on the right-hand sides of operator.cpp:1984-1990 is evaluated in float
					   and only then widened; the 4-term sums of wt_4[] (double array) are double */
					double wqp, wt1, wt2, w4[4];
					wqp = (float)(1 / (LPP[ip] * Cn[SH(nP, 1, n, 0)]) + 1 / (LPP[ip] * Cn[ip]));
					wqp += (float)(1 / (LP[ip] * Cn[SH(nPP, 1, n, 0)]) + 1 / (LP[ip] * Cn[ip]));
					size_t i1 = SH(nP, -1, n, 0); /* Shift(nP,-1) */
					wqp += (float)(1 / (LPP[i1] * Cn[ip]) + 1 / (LPP[i1] * Cn[i1]));
					size_t i2 = SH(nP, -1, nPP, -1); /* Shift(nPP,-1) keeps the nP shift */
					wqp += (float)(1 / (LP[i2] * Cn[i1]) + 1 / (LP[i2] * Cn[i2]));

					w4[0] = (float)(1 / (LPP[ip] * CP[ip]));
					w4[1] = (float)(1 / (LPP[SH(nP, -1, n, 0)] * CP[ip]));
					w4[2] = (float)(1 / (LP[ip] * CPP[ip]));
					w4[3] = (float)(1 / (LP[SH(nPP, -1, n, 0)] * CPP[ip]));
					wt1 = w4[0] + w4[1] + w4[2] + w4[3] - 2 * min4(w4);

					size_t in1 = SH(n, 1, n, 0);
					w4[0] = (float)(1 / (LPP[ip] * CP[in1]));
					w4[1] = (float)(1 / (LPP[SH(nP, -1, n, 0)] * CP[in1]));
					w4[2] = (float)(1 / (LP[ip] * CPP[in1]));
					w4[3] = (float)(1 / (LP[SH(nPP, -1, n, 0)] * CPP[in1]));
					wt2 = w4[0] + w4[1] + w4[2] + w4[3] - 2 * min4(w4);
#undef SH
					double w_total = wqp + wt1 + wt2;
					double newT = 2 / sqrt(w_total);
					if (newT < dT_n && newT > 0.0) dT_n = newT;
				}
		if (dT_n < dT) dT = dT_n;
	}
	return dT;
}

/* Operator::Calc_ECOperatorPos operator.cpp:956-984 */
static void calc_ecop_pos(orc_sim* s, int n, const unsigned pos[3])
{
	size_t nc = ncell(s);
	size_t i = ((size_t)pos[0] * s->N[1] + pos[1]) * s->N[2] + pos[2];
	double dT = s->dT;
	double C = s->EC_C[n * nc + i], G = s->EC_G[n * nc + i];
	size_t o = IDX(s, n, pos[0], pos[1], pos[2]);
	if (C > 0) {
		s->vv[o] = (1.0 - dT * G / 2.0 / C) / (1.0 + dT * G / 2.0 / C);
		s->vi[o] = (dT / C) / (1.0 + dT * G / 2.0 / C);
	} else { s->vv[o] = 0; s->vi[o] = 0; }
	double L = s->EC_L[n * nc + i], R = s->EC_R[n * nc + i];
	if (L > 0) {
		s->ii[o] = (1.0 - dT * R / 2.0 / L) / (1.0 + dT * R / 2.0 / L);
		s->iv[o] = (dT / L) / (1.0 + dT * R / 2.0 / L);
	} else { s->ii[o] = 0; s->iv[o] = 0; }
}

/* Operator::ApplyElectricBC operator.cpp:1099-1138 */
static void apply_electric_bc(orc_sim* s, const int dirs[6])
{
	unsigned pos[3];
	for (int n = 0; n < 3; ++n) {
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		for (pos[nP] = 0; pos[nP] < s->N[nP]; ++pos[nP])
			for (pos[nPP] = 0; pos[nPP] < s->N[nPP]; ++pos[nPP]) {
				if (dirs[2 * n]) {
					pos[n] = 0;
					s->vv[IDX(s, nP, pos[0], pos[1], pos[2])] = 0; s->vi[IDX(s, nP, pos[0], pos[1], pos[2])] = 0;
					s->vv[IDX(s, nPP, pos[0], pos[1], pos[2])] = 0; s->vi[IDX(s, nPP, pos[0], pos[1], pos[2])] = 0;
				}
				if (dirs[2 * n + 1]) {
					pos[n] = s->N[n] - 1;
					for (int c = 0; c < 3; ++c) {
						s->vv[IDX(s, c, pos[0], pos[1], pos[2])] = 0;
						s->vi[IDX(s, c, pos[0], pos[1], pos[2])] = 0;
					}
				}
			}
	}
}

/* Operator::ApplyMagneticBC operator.cpp:1140-1187 */
static void apply_magnetic_bc(orc_sim* s, const int dirs[6])
{
	unsigned pos[3];
	for (int n = 0; n < 3; ++n) {
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		for (pos[nP] = 0; pos[nP] < s->N[nP]; ++pos[nP])
			for (pos[nPP] = 0; pos[nPP] < s->N[nPP]; ++pos[nPP]) {
				if (dirs[2 * n]) {
					pos[n] = 0;
					for (int c = 0; c < 3; ++c) {
						s->ii[IDX(s, c, pos[0], pos[1], pos[2])] = 0;
						s->iv[IDX(s, c, pos[0], pos[1], pos[2])] = 0;
					}
				}
				if (dirs[2 * n + 1]) {
					pos[n] = s->N[n] - 2;
					s->ii[IDX(s, nP, pos[0], pos[1], pos[2])] = 0; s->iv[IDX(s, nP, pos[0], pos[1], pos[2])] = 0;
					s->ii[IDX(s, nPP, pos[0], pos[1], pos[2])] = 0; s->iv[IDX(s, nPP, pos[0], pos[1], pos[2])] = 0;
				}
				pos[n] = s->N[n] - 1;
				for (int c = 0; c < 3; ++c) {
					s->ii[IDX(s, c, pos[0], pos[1], pos[2])] = 0;
					s->iv[IDX(s, c, pos[0], pos[1], pos[2])] = 0;
				}
			}
	}
}

/* Operator::CalcPEC_Range operator.cpp:2046-2084 */
static void calc_pec(orc_sim* s)
{
	int any = 0;
	for (int p = 0; p < s->nprops; ++p) any |= s->props[p].type == P_METAL;
	if (!any) return;
#pragma omp parallel for collapse(2) schedule(static)
	for (unsigned i = 0; i < s->N[0]; ++i)
		for (unsigned j = 0; j < s->N[1]; ++j)
			for (unsigned k = 0; k < s->N[2]; ++k) {
				unsigned pos[3] = {i, j, k};
				double c[3];
				for (int n = 0; n < 3; ++n) {
					yee_coords(s, n, pos, c, 0);
					const prop_t* p = prop_at(s, c, MASK_MAT_METAL);
					if (p && p->type == P_METAL) {
						s->vv[IDX(s, n, i, j, k)] = 0;
						s->vi[IDX(s, n, i, j, k)] = 0;
					}
				}
			}
}

/* Operator::SnapToMeshLine operator.cpp:242-271 (primary mesh, full mesh) */
static unsigned snap_line(const orc_sim* s, int ny, double coord, int* inside)
{
	*inside = 0;
	if (coord < orc_disc_line(s, ny, 0, 0)) return 0;
	unsigned N = s->N[ny];
	if (coord > orc_disc_line(s, ny, N - 1, 0)) return N - 1;
	*inside = 1;
	for (unsigned n = 0; n < N; ++n)
		if (coord <= orc_disc_line(s, ny, n, 1)) return n;
	return 0;
}

/* Operator::Calc_LumpedElements operator.cpp:1586-1763 (parallel RC only; box snapped with
   SnapBox2Mesh operator.cpp:297-344, snap method 0) */
static void calc_lumped(orc_sim* s)
{
	size_t nc = ncell(s);
	for (int l = 0; l < s->nlumped; ++l) {
		lumped_t* le = &s->lumped[l];
		double C = le->C, R = le->R;
		if (C <= 0) C = NAN;
		if (R < 0) R = NAN;
		if (isnan(R) && isnan(C)) continue;
		int ny = le->dir;
		if (ny < 0 || ny > 2) continue;
		int nyP = (ny + 1) % 3, nyPP = (ny + 2) % 3;
		unsigned a[3], b[3];
		int in1, in2, dim = 0, outside = 0;
		for (int n = 0; n < 3; ++n) {
			a[n] = snap_line(s, n, le->start[n], &in1);
			b[n] = snap_line(s, n, le->stop[n], &in2);
			if (!in1 && !in2 && a[n] == b[n]) outside = 1;
			if (b[n] > a[n]) ++dim;
		}
		if (outside || dim <= 0) continue;
		if (a[ny] == b[ny]) continue;
		unsigned pos[3];
		double unitGC = 0;
		for (pos[ny] = a[ny]; pos[ny] < b[ny]; ++pos[ny]) {
			double plane = 0;
			for (pos[nyP] = a[nyP]; pos[nyP] <= b[nyP]; ++pos[nyP])
				for (pos[nyPP] = a[nyPP]; pos[nyPP] <= b[nyPP]; ++pos[nyPP])
					plane += edge_area(s, ny, pos, 0) / orc_edge_length(s, ny, pos, 0);
			unitGC += 1 / plane;
		}
		unitGC = 1 / unitGC;
		double kappa = 0, epsilon = 0;
		if (R > 0) kappa = 1 / R / unitGC;
		if (C > 0) {
			epsilon = C / unitGC;
			if (epsilon < EPS0) C = 0;
		}
		for (pos[ny] = a[ny]; pos[ny] < b[ny]; ++pos[ny])
			for (pos[nyP] = a[nyP]; pos[nyP] <= b[nyP]; ++pos[nyP])
				for (pos[nyPP] = a[nyPP]; pos[nyPP] <= b[nyPP]; ++pos[nyPP]) {
					size_t ip = ((size_t)pos[0] * s->N[1] + pos[1]) * s->N[2] + pos[2];
					if (C > 0) s->EC_C[ny * nc + ip] = epsilon * edge_area(s, ny, pos, 0) / orc_edge_length(s, ny, pos, 0);
					if (R > 0) s->EC_G[ny * nc + ip] = kappa * edge_area(s, ny, pos, 0) / orc_edge_length(s, ny, pos, 0);
					if (R == 0) {
						s->vv[IDX(s, ny, pos[0], pos[1], pos[2])] = 0;
						s->vi[IDX(s, ny, pos[0], pos[1], pos[2])] = 0;
					} else
						calc_ecop_pos(s, ny, pos);
				}
		if (le->caps) {
			for (pos[nyP] = a[nyP]; pos[nyP] <= b[nyP]; ++pos[nyP])
				for (pos[nyPP] = a[nyPP]; pos[nyPP] <= b[nyPP]; ++pos[nyPP]) {
					unsigned ends[2] = {a[ny], b[ny]};
					for (int e = 0; e < 2; ++e) {
						pos[ny] = ends[e];
						if (pos[nyP] < b[nyP]) {
							s->vv[IDX(s, nyP, pos[0], pos[1], pos[2])] = 0;
							s->vi[IDX(s, nyP, pos[0], pos[1], pos[2])] = 0;
						}
						if (pos[nyPP] < b[nyPP]) {
							s->vv[IDX(s, nyPP, pos[0], pos[1], pos[2])] = 0;
							s->vi[IDX(s, nyPP, pos[0], pos[1], pos[2])] = 0;
						}
					}
				}
		}
	}
}

/* ---- excitation extension: Operator_Ext_Excitation::BuildExtension operator_ext_excitation.cpp:105-297 */
typedef struct { unsigned *i[3], *dir, *delay; float* amp; unsigned n, cap; } exc_list;
static void exc_push(exc_list* L, const unsigned pos[3], unsigned dir, float amp, unsigned delay)
{
	if (L->n == L->cap) {
		L->cap = L->cap ? L->cap * 2 : 256;
		for (int n = 0; n < 3; ++n) L->i[n] = realloc(L->i[n], L->cap * sizeof(unsigned));
		L->dir = realloc(L->dir, L->cap * sizeof(unsigned));
		L->delay = realloc(L->delay, L->cap * sizeof(unsigned));
		L->amp = realloc(L->amp, L->cap * sizeof(float));
	}
	for (int n = 0; n < 3; ++n) L->i[n][L->n] = pos[n];
	L->dir[L->n] = dir; L->amp[L->n] = amp; L->delay[L->n] = delay;
	++L->n;
}
static void build_excitation(orc_sim* s)
{
	exc_list V, Cu;
	memset(&V, 0, sizeof(V)); memset(&Cu, 0, sizeof(Cu));
	int any = 0;
	for (int p = 0; p < s->nprops; ++p) any |= s->props[p].type == P_EXCITATION;
	double dT = s->dT;
	unsigned pos[3];
	double c[3];
	if (any)
	for (pos[2] = 0; pos[2] < s->N[2]; ++pos[2])
		for (pos[1] = 0; pos[1] < s->N[1]; ++pos[1])
			for (pos[0] = 0; pos[0] < s->N[0]; ++pos[0]) {
				for (int n = 0; n < 3; ++n) {
					if (!yee_coords(s, n, pos, c, 0)) continue;
					const prop_t* e = prop_at(s, c, MASK_EXC);
					if (!e) continue;
					/* CSPropExcitation::ActiveDir is true for every component unless set otherwise: a hard source
					   zeroes vv/vi of all three components in its box, whatever the excitation vector */
					if (e->exc_type == 0 || e->exc_type == 1) {
						double amp = e->exc_vec[n] * orc_edge_length(s, n, pos, 0);
						if (amp != 0) exc_push(&V, pos, n, (float)amp, (unsigned)(e->delay / dT));
						if (e->exc_type == 1) {
							s->vv[IDX(s, n, pos[0], pos[1], pos[2])] = 0;
							s->vi[IDX(s, n, pos[0], pos[1], pos[2])] = 0;
						}
					}
				}
				for (int n = 0; n < 3; ++n) {
					if (pos[0] >= s->N[0] - 1 || pos[1] >= s->N[1] - 1 || pos[2] >= s->N[2] - 1) continue;
					if (!yee_coords(s, n, pos, c, 1)) continue;
					const prop_t* e = prop_at(s, c, MASK_EXC);
					if (!e) continue;
					if (e->exc_type == 2 || e->exc_type == 3) {
						double amp = e->exc_vec[n] * orc_edge_length(s, n, pos, 1);
						if (amp != 0) exc_push(&Cu, pos, n, (float)amp, (unsigned)(e->delay / dT));
						if (e->exc_type == 3) {
							s->ii[IDX(s, n, pos[0], pos[1], pos[2])] = 0;
							s->iv[IDX(s, n, pos[0], pos[1], pos[2])] = 0;
						}
					}
				}
			}
	s->vcount = V.n; s->ccount = Cu.n;
	for (int n = 0; n < 3; ++n) { s->vidx[n] = V.i[n]; s->cidx[n] = Cu.i[n]; }
	s->vdir = V.dir; s->vdelay = V.delay; s->vamp = V.amp;
	s->cdir = Cu.dir; s->cdelay = Cu.delay; s->camp = Cu.amp;
}

/* ---- Mur: Operator_Ext_Mur_ABC::SetDirection/BuildExtension operator_ext_mur_abc.cpp:80-186 */
static void build_mur(orc_sim* s, mur_t* m, int ny, int top)
{
	memset(m, 0, sizeof(*m));
	m->ny = ny; m->top = top; m->nyP = (ny + 1) % 3; m->nyPP = (ny + 2) % 3;
	if (!top) { m->line = 0; m->shift = 1; }
	else { m->line = s->N[ny] - 1; m->shift = s->N[ny] - 2; }
	m->n[0] = s->N[m->nyP]; m->n[1] = s->N[m->nyPP];
	size_t cnt = (size_t)m->n[0] * m->n[1];
	m->cP = xcalloc(cnt, sizeof(float)); m->cPP = xcalloc(cnt, sizeof(float));
	m->vP = xcalloc(cnt, sizeof(float)); m->vPP = xcalloc(cnt, sizeof(float));
	double dT = s->dT;
	unsigned pos[3] = {0, 0, 0};
	pos[ny] = m->line;
	double delta = fabs(orc_edge_length(s, ny, pos, 0));
	double coord[3];
	if (m->line == 0) coord[ny] = orc_disc_line(s, ny, pos[ny], 0) + delta / 2 / s->grid_delta;
	else coord[ny] = orc_disc_line(s, ny, pos[ny], 0) - delta / 2 / s->grid_delta;
	for (pos[m->nyP] = 0; pos[m->nyP] < m->n[0]; ++pos[m->nyP]) {
		coord[m->nyP] = orc_disc_line(s, m->nyP, pos[m->nyP], 0);
		for (pos[m->nyPP] = 0; pos[m->nyPP] < m->n[1]; ++pos[m->nyPP]) {
			coord[m->nyPP] = orc_disc_line(s, m->nyPP, pos[m->nyPP], 0);
			const prop_t* p = prop_at(s, coord, MASK_MAT);
			double c0t;
			size_t o = (size_t)pos[m->nyP] * m->n[1] + pos[m->nyPP];
			if (p) {
				if (s->mur_vphase > 0.0) c0t = s->mur_vphase * dT;
				else c0t = C0 * dT / sqrt(p->epsR * p->mueR);
				m->cP[o] = (c0t - delta) / (c0t + delta);
				m->cPP[o] = (c0t - delta) / (c0t + delta);
			} else {
				if (s->mur_vphase > 0.0) c0t = s->mur_vphase * dT;
				else c0t = C0 / sqrt(s->bg[0] * s->bg[1]) * dT;
				m->cP[o] = (c0t - delta) / (c0t + delta);
				m->cPP[o] = m->cP[o];
			}
		}
	}
	/* Engine_Ext_Mur_ABC ctor engine_ext_mur_abc.cpp:44-60: delayed start when an excitation
	   sits on the Mur plane */
	int maxDelay = -1;
	for (unsigned n = 0; n < s->vcount; ++n)
		if ((s->vdir[n] == (unsigned)m->nyP || s->vdir[n] == (unsigned)m->nyPP) && s->vidx[ny][n] == m->line)
			if ((int)s->vdelay[n] > maxDelay) maxDelay = (int)s->vdelay[n];
	m->start_ts = 0;
	if (maxDelay >= 0) m->start_ts = maxDelay + s->sig_len + 10;
}

/* ---- UPML: Operator_Ext_UPML::CalcGradingKappa operator_ext_upml.cpp:269-343, default
   grading function :30 evaluated directly (fparser not vendored) */
static double pml_grading(double D, double dl, double W, double Z, double N)
{
	(void)N;
	return -log(1e-6) * log(2.5) / (2 * dl * Z * (pow(2.5, W / dl) - 1)) * pow(2.5, D / dl);
}
static void grading_kappa(const orc_sim* s, int ny, const unsigned pos[3], double Zm, double kv[3], double ki[3])
{
	double depth = 0, width = 0;
	for (int n = 0; n < 3; ++n) {
		unsigned Nn = s->N[n];
		if (pos[n] <= s->pml_size[2 * n] && s->bc[2 * n] == 3) {
			width = (orc_disc_line(s, n, s->pml_size[2 * n], 0) - orc_disc_line(s, n, 0, 0)) * s->grid_delta;
			depth = width - (orc_disc_line(s, n, pos[n], 0) - orc_disc_line(s, n, 0, 0)) * s->grid_delta;
			if (n == ny) depth -= orc_edge_length(s, n, pos, 0) / 2;
			double dl = width / s->pml_size[2 * n], Nv = (double)s->pml_size[2 * n];
			kv[n] = depth > 0 ? pml_grading(depth, dl, width, Zm, Nv) : 0;
			if (n == ny) depth += orc_edge_length(s, n, pos, 0) / 2;
			if (n != ny) depth -= orc_edge_length(s, n, pos, 0) / 2;
			if (depth < 0) depth = 0;
			ki[n] = depth > 0 ? pml_grading(depth, dl, width, Zm, Nv) : 0;
		} else if (pos[n] >= Nn - 1 - s->pml_size[2 * n + 1] && s->bc[2 * n + 1] == 3) {
			width = (orc_disc_line(s, n, Nn - 1, 0) - orc_disc_line(s, n, Nn - s->pml_size[2 * n + 1] - 1, 0)) * s->grid_delta;
			depth = width - (orc_disc_line(s, n, Nn - 1, 0) - orc_disc_line(s, n, pos[n], 0)) * s->grid_delta;
			if (n == ny) depth += orc_edge_length(s, n, pos, 0) / 2;
			/* quirk kept: the upper side uses the LOWER side's size for dl and N (:319) */
			double dl = width / s->pml_size[2 * n], Nv = (double)s->pml_size[2 * n];
			kv[n] = depth > 0 ? pml_grading(depth, dl, width, Zm, Nv) : 0;
			if (n == ny) depth -= orc_edge_length(s, n, pos, 0) / 2;
			if (n != ny) depth += orc_edge_length(s, n, pos, 0) / 2;
			if (depth > width) depth = 0;
			ki[n] = depth > 0 ? pml_grading(depth, dl, width, Zm, Nv) : 0;
		} else { kv[n] = 0; ki[n] = 0; }
	}
}

/* Operator_Ext_UPML::BuildExtension operator_ext_upml.cpp:345-445 */
static void build_upml_box(orc_sim* s, upml_t* u)
{
	size_t cnt = (size_t)3 * u->n[0] * u->n[1] * u->n[2];
	for (int c = 0; c < 6; ++c) u->c[c] = xcalloc(cnt, sizeof(float));
	u->volt_flux = xcalloc(cnt, sizeof(float));
	u->curr_flux = xcalloc(cnt, sizeof(float));
	double dT = s->dT;
#pragma omp parallel for collapse(2) schedule(static)
	for (unsigned li = 0; li < u->n[0]; ++li)
		for (unsigned lj = 0; lj < u->n[1]; ++lj)
			for (unsigned lk = 0; lk < u->n[2]; ++lk) {
				unsigned pos[3] = {li + u->start[0], lj + u->start[1], lk + u->start[2]};
				for (int n = 0; n < 3; ++n) {
					double em[4], kv[3] = {0, 0, 0}, ki[3] = {0, 0, 0};
					eff_mat(s, n, pos, em);
					grading_kappa(s, n, pos, Z0, kv, ki);
					int nP = (n + 1) % 3, nPP = (n + 2) % 3;
					size_t lo = (((size_t)n * u->n[0] + li) * u->n[1] + lj) * u->n[2] + lk;
					size_t go = IDX(s, n, pos[0], pos[1], pos[2]);
					if ((kv[0] + kv[1] + kv[2]) != 0 && em[1] < 1e3) {
						if ((s->vv[go] + s->vi[go]) != 0) {
							s->vv[go] = (2 * EPS0 - kv[nP] * dT) / (2 * EPS0 + kv[nP] * dT);
							s->vi[go] = (2 * EPS0 * dT) / (2 * EPS0 + kv[nP] * dT) * orc_edge_length(s, n, pos, 0) / edge_area(s, n, pos, 0);
							u->c[0][lo] = (2 * EPS0 - kv[nPP] * dT) / (2 * EPS0 + kv[nPP] * dT);
							u->c[1][lo] = (2 * EPS0 + kv[n] * dT) / (2 * EPS0 + kv[nPP] * dT) / em[0];
							u->c[2][lo] = (2 * EPS0 - kv[n] * dT) / (2 * EPS0 + kv[nPP] * dT) / em[0];
						}
					} else {
						u->c[0][lo] = s->vv[go];
						s->vv[go] = 0;
						u->c[2][lo] = 0;
						u->c[1][lo] = 1;
					}
					if ((ki[0] + ki[1] + ki[2]) != 0) {
						if ((s->ii[go] + s->iv[go]) != 0) {
							s->ii[go] = (2 * EPS0 - ki[nP] * dT) / (2 * EPS0 + ki[nP] * dT);
							s->iv[go] = (2 * EPS0 * dT) / (2 * EPS0 + ki[nP] * dT) * orc_edge_length(s, n, pos, 1) / edge_area(s, n, pos, 1);
							u->c[3][lo] = (2 * EPS0 - ki[nPP] * dT) / (2 * EPS0 + ki[nPP] * dT);
							u->c[4][lo] = (2 * EPS0 + ki[n] * dT) / (2 * EPS0 + ki[nPP] * dT) / em[2];
							u->c[5][lo] = (2 * EPS0 - ki[n] * dT) / (2 * EPS0 + ki[nPP] * dT) / em[2];
						}
					} else {
						u->c[3][lo] = s->ii[go];
						s->ii[go] = 0;
						u->c[5][lo] = 0;
						u->c[4][lo] = 1;
					}
				}
			}
}

/* Operator_Ext_UPML::Create_UPML operator_ext_upml.cpp:69-247 (Cartesian part) */
static void create_upml(orc_sim* s)
{
	int BC[6]; unsigned size[6];
	for (int n = 0; n < 6; ++n) { BC[n] = s->bc[n]; size[n] = s->pml_size[n]; }
	for (int n = 0; n < 3; ++n)
		if ((size[2 * n] * (BC[2 * n] == 3) + size[2 * n + 1] * (BC[2 * n + 1] == 3)) >= s->N[n]) {
			fprintf(stderr, "oracle: not enough lines for the pml in direction %d, resetting to PEC\n", n);
			BC[2 * n] = 0; size[2 * n] = 0; BC[2 * n + 1] = 0; size[2 * n + 1] = 0;
		}
	/* note: like the reference the boxes keep using s->bc / s->pml_size inside CalcGradingKappa
	   (SetBoundaryCondition copies the adjusted arrays) */
	for (int n = 0; n < 6; ++n) { s->bc[n] = BC[n] == 3 ? 3 : (s->bc[n] == 3 ? 0 : s->bc[n]); s->pml_size[n] = size[n]; }
	unsigned start[3] = {0, 0, 0};
	unsigned stop[3] = {s->N[0] - 1, s->N[1] - 1, s->N[2] - 1};
	s->nupml = 0;
#define ADD_BOX() do { upml_t* u = &s->upml[s->nupml++]; memset(u, 0, sizeof(*u)); \
	for (int q = 0; q < 3; ++q) { u->start[q] = start[q]; u->n[q] = stop[q] - start[q] + 1; } } while (0)
	if (BC[0] == 3) { start[0] = 0; stop[0] = size[0]; ADD_BOX(); }
	if (BC[1] == 3) { start[0] = s->N[0] - 1 - size[1]; stop[0] = s->N[0] - 1; ADD_BOX(); }
	start[0] = (size[0] + 1) * (BC[0] == 3);
	stop[0] = s->N[0] - 1 - (size[0] + 1) * (BC[1] == 3); /* reference uses size[0] here (:164) */
	if (BC[2] == 3) { start[1] = 0; stop[1] = size[2]; ADD_BOX(); }
	if (BC[3] == 3) { start[1] = s->N[1] - 1 - size[3]; stop[1] = s->N[1] - 1; ADD_BOX(); }
	start[1] = (size[2] + 1) * (BC[2] == 3);
	stop[1] = s->N[1] - 1 - (size[3] + 1) * (BC[3] == 3);
	if (BC[4] == 3) { start[2] = 0; stop[2] = size[4]; ADD_BOX(); }
	if (BC[5] == 3) { start[2] = s->N[2] - 1 - size[5]; stop[2] = s->N[2] - 1; ADD_BOX(); }
#undef ADD_BOX
}

/* ---- Lorentz/Drude: Operator_Ext_LorentzMaterial::BuildExtension operator_ext_lorentzmaterial.cpp:120-445
   (Lorentz material part; Debye not restated) */
typedef struct { double* d; unsigned n, cap; } dvec;
static void dpush(dvec* v, double x)
{
	if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->d = realloc(v->d, v->cap * sizeof(double)); }
	v->d[v->n++] = x;
}
static void build_lorentz(orc_sim* s)
{
	s->lor_order = 0;
	for (int p = 0; p < s->nprops; ++p)
		if (s->props[p].type == P_LORENTZ && s->props[p].order > s->lor_order) s->lor_order = s->props[p].order;
	double dT = s->dT;
	size_t nc = ncell(s);
	for (int order = 0; order < s->lor_order; ++order) {
		lor_order_t* L = &s->lor[order];
		memset(L, 0, sizeof(*L));
		dvec vpos[3] = {{0}}, v_int[3] = {{0}}, v_ext[3] = {{0}}, i_int[3] = {{0}}, i_ext[3] = {{0}}, v_Lor[3] = {{0}}, i_Lor[3] = {{0}};
		unsigned pos[3];
		double coord[3];
		for (pos[0] = 0; pos[0] < s->N[0]; ++pos[0])
			for (pos[1] = 0; pos[1] < s->N[1]; ++pos[1])
				for (pos[2] = 0; pos[2] < s->N[2]; ++pos[2]) {
					size_t index = ((size_t)pos[0] * s->N[1] + pos[1]) * s->N[2] + pos[2];
					int b_pos_on = 0;
					double L_D[3], R_D[3], C_L[3], C_D[3], G_D[3], L_L[3];
					for (int n = 0; n < 3; ++n) {
						L_D[n] = 0; R_D[n] = 0; C_L[n] = 0;
						if (!yee_coords(s, n, pos, coord, 0)) continue;
						if (s->vi[IDX(s, n, pos[0], pos[1], pos[2])] == 0) continue;
						const prop_t* p = prop_at(s, coord, MASK_MAT_METAL);
						if (!p || p->type != P_LORENTZ) continue;
						double w_plasma = (order < p->order ? p->eps_fp[order] : 0) * 2 * ORC_PI;
						if (w_plasma > 0 && s->EC_C[n * nc + index] > 0) {
							b_pos_on = 1; L->volt_on = 1;
							L_D[n] = 1 / (w_plasma * w_plasma * s->EC_C[n * nc + index]);
						}
						double t_relax = order < p->order ? p->eps_tau[order] : 0;
						if (t_relax > 0 && L->volt_on) R_D[n] = L_D[n] / t_relax;
						double w_Lor = (order < p->order ? p->eps_flor[order] : 0) * 2 * ORC_PI;
						if (w_Lor > 0 && L_D[n] > 0) { L->volt_lor_on = 1; C_L[n] = 1 / (w_Lor * w_Lor * L_D[n]); }
					}
					for (int n = 0; n < 3; ++n) {
						C_D[n] = 0; G_D[n] = 0; L_L[n] = 0;
						if (!yee_coords(s, n, pos, coord, 1)) continue;
						if (s->iv[IDX(s, n, pos[0], pos[1], pos[2])] == 0) continue;
						const prop_t* p = prop_at(s, coord, MASK_MAT_METAL);
						if (!p || p->type != P_LORENTZ) continue;
						double w_plasma = (order < p->order ? p->mue_fp[order] : 0) * 2 * ORC_PI;
						if (w_plasma > 0 && s->EC_L[n * nc + index] > 0) {
							b_pos_on = 1; L->curr_on = 1;
							C_D[n] = 1 / (w_plasma * w_plasma * s->EC_L[n * nc + index]);
						}
						double t_relax = order < p->order ? p->mue_tau[order] : 0;
						if (t_relax > 0 && L->curr_on) G_D[n] = C_D[n] / t_relax;
						double w_Lor = (order < p->order ? p->mue_flor[order] : 0) * 2 * ORC_PI;
						if (w_Lor > 0 && C_D[n] > 0) { L->curr_lor_on = 1; L_L[n] = 1 / (w_Lor * w_Lor * C_D[n]); }
					}
					if (!b_pos_on) continue;
					for (int n = 0; n < 3; ++n) {
						double VI = s->vi[IDX(s, n, pos[0], pos[1], pos[2])];
						double IV = s->iv[IDX(s, n, pos[0], pos[1], pos[2])];
						dpush(&vpos[n], pos[n]);
						if (L_D[n] > 0) {
							dpush(&v_int[n], (2.0 * L_D[n] - dT * R_D[n]) / (2.0 * L_D[n] + dT * R_D[n]));
							dpush(&v_ext[n], dT / (L_D[n] + dT * R_D[n] / 2.0) * VI);
						} else if (R_D[n] > 0 && C_L[n] > 0) {
							dpush(&v_int[n], (2.0 * dT - R_D[n] * C_L[n]) / (C_L[n] * R_D[n]));
							dpush(&v_ext[n], 2.0 / R_D[n] * VI);
						} else { dpush(&v_int[n], 1); dpush(&v_ext[n], 0); }
						if (C_D[n] > 0) {
							dpush(&i_int[n], (2.0 * C_D[n] - dT * G_D[n]) / (2.0 * C_D[n] + dT * G_D[n]));
							dpush(&i_ext[n], dT / (C_D[n] + dT * G_D[n] / 2.0) * IV);
						} else { dpush(&i_int[n], 1); dpush(&i_ext[n], 0); }
						dpush(&v_Lor[n], C_L[n] > 0 ? dT / C_L[n] / VI : 0);
						dpush(&i_Lor[n], L_L[n] > 0 ? dT / L_L[n] / IV : 0);
					}
				}
		L->count = vpos[0].n;
		for (int n = 0; n < 3; ++n) {
			unsigned cnt = L->count;
			L->pos[n] = xcalloc(cnt, sizeof(unsigned));
			for (unsigned i = 0; i < cnt; ++i) L->pos[n][i] = (unsigned)vpos[n].d[i];
#define CP(dst, src) do { dst = xcalloc(cnt, sizeof(float)); for (unsigned i = 0; i < cnt; ++i) dst[i] = (float)src.d[i]; } while (0)
			if (L->volt_on) { CP(L->v_int[n], v_int[n]); CP(L->v_ext[n], v_ext[n]); L->volt_ADE[n] = xcalloc(cnt, sizeof(float)); }
			if (L->curr_on) { CP(L->i_int[n], i_int[n]); CP(L->i_ext[n], i_ext[n]); L->curr_ADE[n] = xcalloc(cnt, sizeof(float)); }
			if (L->volt_lor_on) { CP(L->v_lor[n], v_Lor[n]); L->volt_Lor_ADE[n] = xcalloc(cnt, sizeof(float)); }
			if (L->curr_lor_on) { CP(L->i_lor[n], i_Lor[n]); L->curr_Lor_ADE[n] = xcalloc(cnt, sizeof(float)); }
#undef CP
			free(vpos[n].d); free(v_int[n].d); free(v_ext[n].d); free(i_int[n].d); free(i_ext[n].d);
			free(v_Lor[n].d); free(i_Lor[n].d);
		}
	}
}

/* ---------------------------------------------------------------- engine extensions */
#define VOLT(s, n, p) (*orc_vref(s, n, p))
#define CURR(s, n, p) (*orc_cref(s, n, p))

/* Engine_Ext_Excitation::Apply2Voltages engine_ext_excitation.cpp:33-60 */
static void exc_applyV(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	int numTS = (int)s->numTS;
	unsigned length = s->sig_len;
	int p = numTS + 1;
	if (s->exc_period > 0) p = (int)(s->exc_period / s->dT);
	for (unsigned n = 0; n < s->vcount; ++n) {
		int exc_pos = numTS - (int)s->vdelay[n];
		exc_pos *= (exc_pos > 0);
		exc_pos %= p;
		exc_pos *= (exc_pos < (int)length);
		unsigned pos[3] = {s->vidx[0][n], s->vidx[1][n], s->vidx[2][n]};
		VOLT(s, s->vdir[n], pos) = VOLT(s, s->vdir[n], pos) + s->vamp[n] * s->sig_v[exc_pos];
	}
}
/* Engine_Ext_Excitation::Apply2Current engine_ext_excitation.cpp:67-94 */
static void exc_applyI(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	int numTS = (int)s->numTS;
	unsigned length = s->sig_len;
	int p = numTS + 1;
	if (s->exc_period > 0) p = (int)(s->exc_period / s->dT);
	for (unsigned n = 0; n < s->ccount; ++n) {
		int exc_pos = numTS - (int)s->cdelay[n];
		exc_pos *= (exc_pos > 0);
		exc_pos %= p;
		exc_pos *= (exc_pos < (int)length);
		unsigned pos[3] = {s->cidx[0][n], s->cidx[1][n], s->cidx[2][n]};
		CURR(s, s->cdir[n], pos) = CURR(s, s->cdir[n], pos) + s->camp[n] * s->sig_i[exc_pos];
	}
}

/* Engine_Ext_Mur_ABC engine_ext_mur_abc.cpp:82-173 */
static void mur_preV(orc_sim* s, ext_t* e, int tid, int nth)
{
	mur_t* m = e->data;
	if (s->numTS < m->start_ts) return;
	unsigned pos[3] = {0, 0, 0}, ps[3] = {0, 0, 0};
	pos[m->ny] = m->line; ps[m->ny] = m->shift;
	unsigned x0, xn;
	orc_jobs(m->n[0], nth, tid, &x0, &xn); /* engine_ext_mur_abc.cpp:68-77 */
	for (unsigned i = x0; i < x0 + xn; ++i) {
		pos[m->nyP] = i; ps[m->nyP] = i;
		for (unsigned j = 0; j < m->n[1]; ++j) {
			pos[m->nyPP] = j; ps[m->nyPP] = j;
			size_t o = (size_t)i * m->n[1] + j;
			m->vP[o] = VOLT(s, m->nyP, ps) - m->cP[o] * VOLT(s, m->nyP, pos);
			m->vPP[o] = VOLT(s, m->nyPP, ps) - m->cPP[o] * VOLT(s, m->nyPP, pos);
		}
	}
}
static void mur_postV(orc_sim* s, ext_t* e, int tid, int nth)
{
	mur_t* m = e->data;
	if (s->numTS < m->start_ts) return;
	unsigned ps[3] = {0, 0, 0};
	ps[m->ny] = m->shift;
	unsigned x0, xn;
	orc_jobs(m->n[0], nth, tid, &x0, &xn);
	for (unsigned i = x0; i < x0 + xn; ++i) {
		ps[m->nyP] = i;
		for (unsigned j = 0; j < m->n[1]; ++j) {
			ps[m->nyPP] = j;
			size_t o = (size_t)i * m->n[1] + j;
			m->vP[o] += m->cP[o] * VOLT(s, m->nyP, ps);
			m->vPP[o] += m->cPP[o] * VOLT(s, m->nyPP, ps);
		}
	}
}
static void mur_applyV(orc_sim* s, ext_t* e, int tid, int nth)
{
	mur_t* m = e->data;
	if (s->numTS < m->start_ts) return;
	unsigned pos[3] = {0, 0, 0};
	pos[m->ny] = m->line;
	unsigned x0, xn;
	orc_jobs(m->n[0], nth, tid, &x0, &xn);
	for (unsigned i = x0; i < x0 + xn; ++i) {
		pos[m->nyP] = i;
		for (unsigned j = 0; j < m->n[1]; ++j) {
			pos[m->nyPP] = j;
			size_t o = (size_t)i * m->n[1] + j;
			VOLT(s, m->nyP, pos) = m->vP[o];
			VOLT(s, m->nyPP, pos) = m->vPP[o];
		}
	}
}

/* Engine_Ext_UPML engine_ext_upml.cpp:52-229 */
static void upml_preV(orc_sim* s, ext_t* e, int tid, int nth)
{
	upml_t* u = e->data;
	unsigned x0, xn;
	orc_jobs(u->n[0], nth, tid, &x0, &xn); /* engine_ext_upml.cpp:43-50 */
	for (unsigned li = x0; li < x0 + xn; ++li)
		for (unsigned lj = 0; lj < u->n[1]; ++lj)
			for (unsigned lk = 0; lk < u->n[2]; ++lk) {
				unsigned pos[3] = {li + u->start[0], lj + u->start[1], lk + u->start[2]};
				for (int n = 0; n < 3; ++n) {
					size_t lo = (((size_t)n * u->n[0] + li) * u->n[1] + lj) * u->n[2] + lk;
					float f = u->c[0][lo] * VOLT(s, n, pos) - u->c[2][lo] * u->volt_flux[lo];
					VOLT(s, n, pos) = u->volt_flux[lo];
					u->volt_flux[lo] = f;
				}
			}
}
static void upml_postV(orc_sim* s, ext_t* e, int tid, int nth)
{
	upml_t* u = e->data;
	unsigned x0, xn;
	orc_jobs(u->n[0], nth, tid, &x0, &xn); /* engine_ext_upml.cpp:43-50 */
	for (unsigned li = x0; li < x0 + xn; ++li)
		for (unsigned lj = 0; lj < u->n[1]; ++lj)
			for (unsigned lk = 0; lk < u->n[2]; ++lk) {
				unsigned pos[3] = {li + u->start[0], lj + u->start[1], lk + u->start[2]};
				for (int n = 0; n < 3; ++n) {
					size_t lo = (((size_t)n * u->n[0] + li) * u->n[1] + lj) * u->n[2] + lk;
					float f = u->volt_flux[lo];
					u->volt_flux[lo] = VOLT(s, n, pos);
					VOLT(s, n, pos) = f + u->c[1][lo] * u->volt_flux[lo];
				}
			}
}
static void upml_preI(orc_sim* s, ext_t* e, int tid, int nth)
{
	upml_t* u = e->data;
	unsigned x0, xn;
	orc_jobs(u->n[0], nth, tid, &x0, &xn); /* engine_ext_upml.cpp:43-50 */
	for (unsigned li = x0; li < x0 + xn; ++li)
		for (unsigned lj = 0; lj < u->n[1]; ++lj)
			for (unsigned lk = 0; lk < u->n[2]; ++lk) {
				unsigned pos[3] = {li + u->start[0], lj + u->start[1], lk + u->start[2]};
				for (int n = 0; n < 3; ++n) {
					size_t lo = (((size_t)n * u->n[0] + li) * u->n[1] + lj) * u->n[2] + lk;
					float f = u->c[3][lo] * CURR(s, n, pos) - u->c[5][lo] * u->curr_flux[lo];
					CURR(s, n, pos) = u->curr_flux[lo];
					u->curr_flux[lo] = f;
				}
			}
}
static void upml_postI(orc_sim* s, ext_t* e, int tid, int nth)
{
	upml_t* u = e->data;
	unsigned x0, xn;
	orc_jobs(u->n[0], nth, tid, &x0, &xn); /* engine_ext_upml.cpp:43-50 */
	for (unsigned li = x0; li < x0 + xn; ++li)
		for (unsigned lj = 0; lj < u->n[1]; ++lj)
			for (unsigned lk = 0; lk < u->n[2]; ++lk) {
				unsigned pos[3] = {li + u->start[0], lj + u->start[1], lk + u->start[2]};
				for (int n = 0; n < 3; ++n) {
					size_t lo = (((size_t)n * u->n[0] + li) * u->n[1] + lj) * u->n[2] + lk;
					float f = u->curr_flux[lo];
					u->curr_flux[lo] = CURR(s, n, pos);
					CURR(s, n, pos) = f + u->c[4][lo] * u->curr_flux[lo];
				}
			}
}

/* Engine_Ext_LorentzMaterial::DoPreVoltageUpdates engine_ext_lorentzmaterial.cpp:79-120 */
static void lor_preV(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	for (int o = 0; o < s->lor_order; ++o) {
		lor_order_t* L = &s->lor[o];
		if (!L->volt_on) continue;
		for (unsigned i = 0; i < L->count; ++i) {
			unsigned pos[3] = {L->pos[0][i], L->pos[1][i], L->pos[2][i]};
			for (int n = 0; n < 3; ++n) {
				if (L->volt_lor_on) {
					L->volt_Lor_ADE[n][i] += L->v_lor[n][i] * L->volt_ADE[n][i];
					L->volt_ADE[n][i] *= L->v_int[n][i];
					L->volt_ADE[n][i] += L->v_ext[n][i] * (VOLT(s, n, pos) - L->volt_Lor_ADE[n][i]);
				} else {
					L->volt_ADE[n][i] *= L->v_int[n][i];
					L->volt_ADE[n][i] += L->v_ext[n][i] * VOLT(s, n, pos);
				}
			}
		}
	}
}
/* engine_ext_lorentzmaterial.cpp:127-168 */
static void lor_preI(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	for (int o = 0; o < s->lor_order; ++o) {
		lor_order_t* L = &s->lor[o];
		if (!L->curr_on) continue;
		for (unsigned i = 0; i < L->count; ++i) {
			unsigned pos[3] = {L->pos[0][i], L->pos[1][i], L->pos[2][i]};
			for (int n = 0; n < 3; ++n) {
				if (L->curr_lor_on) {
					L->curr_Lor_ADE[n][i] += L->i_lor[n][i] * L->curr_ADE[n][i];
					L->curr_ADE[n][i] *= L->i_int[n][i];
					L->curr_ADE[n][i] += L->i_ext[n][i] * (CURR(s, n, pos) - L->curr_Lor_ADE[n][i]);
				} else {
					L->curr_ADE[n][i] *= L->i_int[n][i];
					L->curr_ADE[n][i] += L->i_ext[n][i] * CURR(s, n, pos);
				}
			}
		}
	}
}
/* Engine_Ext_Dispersive::Apply2Voltages / Apply2Current engine_ext_dispersive.cpp:76-127 */
static void lor_applyV(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	for (int o = 0; o < s->lor_order; ++o) {
		lor_order_t* L = &s->lor[o];
		if (!L->volt_on) continue;
		for (unsigned i = 0; i < L->count; ++i) {
			unsigned pos[3] = {L->pos[0][i], L->pos[1][i], L->pos[2][i]};
			for (int n = 0; n < 3; ++n) VOLT(s, n, pos) = VOLT(s, n, pos) - L->volt_ADE[n][i];
		}
	}
}
static void lor_applyI(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	for (int o = 0; o < s->lor_order; ++o) {
		lor_order_t* L = &s->lor[o];
		if (!L->curr_on) continue;
		for (unsigned i = 0; i < L->count; ++i) {
			unsigned pos[3] = {L->pos[0][i], L->pos[1][i], L->pos[2][i]};
			for (int n = 0; n < 3; ++n) CURR(s, n, pos) = CURR(s, n, pos) - L->curr_ADE[n][i];
		}
	}
}

/* Engine_Ext_LumpedRLC engine_ext_lumpedRLC.cpp:83-142 */
static void rlc_preV(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	(void)s;
	rlc_t* R = e->data;
	float* t = R->Vdn[2]; R->Vdn[2] = R->Vdn[1]; R->Vdn[1] = R->Vdn[0]; R->Vdn[0] = t;
	for (unsigned p = 0; p < R->count; ++p) R->Il[p] += (R->i2v[p]) * (R->ilv[p]) * R->Vdn[1][p];
}
static void rlc_applyV(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return; /* no threadID overload: thread 0 only, engine_extension.cpp:64-69 */
	(void)nth;
	rlc_t* R = e->data;
	float* t = R->Jn[2]; R->Jn[2] = R->Jn[1]; R->Jn[1] = R->Jn[0]; R->Jn[0] = t;
	for (unsigned p = 0; p < R->count; ++p) {
		unsigned pos[3] = {R->pos[0][p], R->pos[1][p], R->pos[2][p]};
		R->Vdn[0][p] = VOLT(s, R->dir[p], pos);
	}
	for (unsigned p = 0; p < R->count; ++p) {
		R->Vdn[0][p] = (R->vvd[p]) * (R->Vdn[0][p] - R->Il[p] + (R->vv2[p]) * R->Vdn[2][p] +
		                              (R->vj1[p]) * R->Jn[1][p] + (R->vj2[p]) * R->Jn[2][p]);
		R->Jn[0][p] = (R->ib0[p]) * (R->Vdn[0][p] - R->Vdn[2][p]) -
		              ((R->b1[p]) * (R->ib0[p])) * R->Jn[1][p] - ((R->b2[p]) * (R->ib0[p])) * R->Jn[2][p];
	}
	for (unsigned p = 0; p < R->count; ++p) {
		unsigned pos[3] = {R->pos[0][p], R->pos[1][p], R->pos[2][p]};
		VOLT(s, R->dir[p], pos) = R->Vdn[0][p];
	}
}

/* Engine_Ext_SteadyState::Apply2Voltages engine_ext_steadystate.cpp:50-107 */
static void ss_applyV(orc_sim* s, ext_t* e, int tid, int nth)
{
	if (tid != 0) return;
	(void)nth;
	ss_t* S = e->data;
	const unsigned p = S->period, TS = s->numTS, rel_pos = TS % (2 * p);
	for (unsigned n = 0; n < S->count; ++n) {
		unsigned pos[3] = {S->pos[0][n], S->pos[1][n], S->pos[2][n]};
		S->rec[(size_t)n * 2 * p + rel_pos] = VOLT(s, S->dir[n], pos);
	}
	if ((TS % p == 0) && (TS >= 2 * p)) {
		int no_valid = 1;
		S->last_max_diff = 0;
		double curr_total_energy = orc_energy(s);
		if (S->last_total_energy > 0) {
			S->last_max_diff = fabs(curr_total_energy - S->last_total_energy) / S->last_total_energy;
			no_valid = 0;
		}
		S->last_total_energy = curr_total_energy;
		unsigned old_pos = 0, new_pos = p;
		if (rel_pos <= p) { new_pos = 0; old_pos = p; }
		double max_pow = 0;
		double* curr_pow = xcalloc(S->count, sizeof(double));
		double* diff_pow = xcalloc(S->count, sizeof(double));
		for (unsigned n = 0; n < S->count; ++n) {
			const double* buf = S->rec + (size_t)n * 2 * p;
			for (unsigned nt = 0; nt < p; ++nt) {
				curr_pow[n] += buf[nt + new_pos] * buf[nt + new_pos];
				diff_pow[n] += (buf[nt + old_pos] - buf[nt + new_pos]) * (buf[nt + old_pos] - buf[nt + new_pos]);
			}
			if (curr_pow[n] > max_pow) max_pow = curr_pow[n];
		}
		for (unsigned n = 0; n < S->count; ++n)
			if (curr_pow[n] > max_pow * 1e-2) {
				if (diff_pow[n] / curr_pow[n] > S->last_max_diff) S->last_max_diff = diff_pow[n] / curr_pow[n];
				no_valid = 0;
			}
		if (no_valid || S->last_max_diff > 1) S->last_max_diff = 1;
		free(curr_pow); free(diff_pow);
	}
}

int orc_add_steadystate(orc_sim* s, unsigned period_ts, unsigned count, const unsigned* pos3, const int* dir)
{
	if (s->ss || period_ts == 0) return -1;
	ss_t* S = xcalloc(1, sizeof(ss_t));
	S->period = period_ts; S->count = count;
	S->dir = xcalloc(count, sizeof(int));
	memcpy(S->dir, dir, count * sizeof(int));
	for (int n = 0; n < 3; ++n) {
		S->pos[n] = xcalloc(count, sizeof(unsigned));
		memcpy(S->pos[n], pos3 + (size_t)n * count, count * sizeof(unsigned));
	}
	S->rec = xcalloc((size_t)count * 2 * period_ts, sizeof(double));
	S->last_max_diff = 1; S->last_total_energy = 0;
	s->ss = S;
	return 0;
}
double orc_steadystate_last_diff(const orc_sim* s) { return s->ss ? s->ss->last_max_diff : 1.0; }

static void add_ext(orc_sim* s, int prio, void* data, hook_fn preV, hook_fn postV, hook_fn applyV,
                    hook_fn preI, hook_fn postI, hook_fn applyI)
{
	ext_t* e = &s->exts[s->nexts++];
	e->prio = prio; e->data = data;
	e->preV = preV; e->postV = postV; e->applyV = applyV;
	e->preI = preI; e->postI = postI; e->applyI = applyI;
}

/* ---- TFSF plane wave ---------------------------------------------------------------------
   Operator_Ext_TFSF::BuildExtension operator_ext_tfsf.cpp:86-406 for ONE plane-wave excitation
   (excite type 10) on the box of mesh indices start..stop; Engine_Ext_TFSF engine_ext_tfsf.cpp:36-215.
   frequency <= 0 only: the phase velocity is c0/n (line 160-161); the numeric phase velocity of
   line 163 (Operator::CalcNumericPhaseVelocity) is not restated. */
int orc_set_tfsf(orc_sim* s, const unsigned start[3], const unsigned stop[3], const double prop_dir[3], const double e_amp[3])
{
	if (s->built) return -1;
	tfsf_t* t = &s->tfsf;
	memset(t, 0, sizeof(*t));
	for (int n = 0; n < 3; ++n) {
		if (start[n] > stop[n] || stop[n] >= s->N[n]) return -2;
		t->start[n] = start[n]; t->stop[n] = stop[n];
		t->prop_dir[n] = prop_dir[n]; t->e_amp[n] = e_amp[n];
	}
	t->on = 1;
	return 0;
}
#define PW_DIST(c, o, d) (fabs(((c)[0] - (o)[0]) * (d)[0]) + fabs(((c)[1] - (o)[1]) * (d)[1]) + fabs(((c)[2] - (o)[2]) * (d)[2]))
static int build_tfsf(orc_sim* s)
{
	tfsf_t* t = &s->tfsf;
	const double dT = s->dT;
	const double ref_index = sqrt(s->bg[0] * s->bg[1]);
	double dir_norm = sqrt(t->prop_dir[0] * t->prop_dir[0] + t->prop_dir[1] * t->prop_dir[1] + t->prop_dir[2] * t->prop_dir[2]);
	if (dir_norm == 0) { t->on = 0; return 0; }                                     /* :134-139 */
	for (int n = 0; n < 3; ++n) t->prop_dir[n] /= dir_norm;
	t->ph_vel = C0 / ref_index;                                                       /* :160-161 */
	double origin[3];
	int inc_low[3];
	for (int n = 0; n < 3; ++n) {                                                     /* :173-197 */
		t->nl[n] = t->stop[n] - t->start[n] + 1;
		inc_low[n] = t->prop_dir[n] >= 0;
		t->active[n][0] = t->start[n] != 0;
		t->active[n][1] = t->stop[n] != s->N[n] - 1;
		unsigned ui = inc_low[n] ? t->start[n] - 1 : t->stop[n] + 1;
		origin[n] = orc_disc_line(s, n, ui, 0);
	}
	double* E = t->e_amp;
	double dotEk = E[0] * t->prop_dir[0] + E[1] * t->prop_dir[1] + E[2] * t->prop_dir[2];
	double angle = acos(dotEk / (E[0] * E[0] + E[1] * E[1] + E[2] * E[2])) / M_PI * 180;
	if (angle == 0) { t->on = 0; return 0; }                                          /* :202-207 */
	if (angle != 90)
		for (int n = 0; n < 3; ++n) E[n] -= t->prop_dir[n] * dotEk;                   /* :208-214 */
	for (int n = 0; n < 3; ++n) {                                                     /* :217-223 */
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		t->h_amp[n] = t->prop_dir[nP] * E[nPP] - t->prop_dir[nPP] * E[nP];
		t->h_amp[n] /= Z0 * sqrt(s->bg[1] / s->bg[0]);
	}
	const double unit = s->grid_delta;
	unsigned max_delay = 0;
	double coord[3], dist, delay;
	unsigned pos[3];
	for (int n = 0; n < 3; ++n) {
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		pos[n] = 0;
		pos[nP] = t->start[nP];
		unsigned numP = t->nl[nP] * t->nl[nPP];
		if (!t->active[n][0] && !t->active[n][1]) continue;
		for (int l = 0; l < 2; ++l)
			for (int c = 0; c < 2; ++c)
				if (t->active[n][l]) {
					t->vdelay[n][l][c] = xcalloc(numP, sizeof(unsigned)); t->vdd[n][l][c] = xcalloc(numP, sizeof(float)); t->vamp[n][l][c] = xcalloc(numP, sizeof(float));
					t->cdelay[n][l][c] = xcalloc(numP, sizeof(unsigned)); t->cdd[n][l][c] = xcalloc(numP, sizeof(float)); t->camp[n][l][c] = xcalloc(numP, sizeof(float));
				}
		unsigned ui_pos = 0;
		for (unsigned i = 0; i < t->nl[nP]; ++i) {
			pos[nPP] = t->start[nPP];
			for (unsigned j = 0; j < t->nl[nPP]; ++j) {
#define TF_SET(DEL, DD, AMP, l, c, extra, ampval) do { \
	delay = dist * unit / t->ph_vel / dT + (extra); \
	if ((unsigned)delay > max_delay) max_delay = (unsigned)delay; \
	t->DEL[n][l][c][ui_pos] = (unsigned)floor(delay); \
	t->DD[n][l][c][ui_pos] = (float)(delay - floor(delay)); \
	t->AMP[n][l][c][ui_pos] = (float)(ampval); } while (0)
				/* current updates :257-305 */
				pos[n] = t->start[n];
				if (t->active[n][0]) {
					yee_coords(s, nP, pos, coord, 0); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(cdelay, cdd, camp, 0, 1, 0.0, E[nP] * orc_edge_length(s, nP, pos, 0));
					yee_coords(s, nPP, pos, coord, 0); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(cdelay, cdd, camp, 0, 0, 0.0, E[nPP] * orc_edge_length(s, nPP, pos, 0));
					--pos[n];
					t->camp[n][0][0][ui_pos] *= s->iv[IDX(s, nP, pos[0], pos[1], pos[2])];
					t->camp[n][0][1][ui_pos] *= s->iv[IDX(s, nPP, pos[0], pos[1], pos[2])];
				}
				if (t->active[n][1]) {
					pos[n] = t->stop[n];
					yee_coords(s, nP, pos, coord, 0); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(cdelay, cdd, camp, 1, 1, 0.0, E[nP] * orc_edge_length(s, nP, pos, 0));
					yee_coords(s, nPP, pos, coord, 0); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(cdelay, cdd, camp, 1, 0, 0.0, E[nPP] * orc_edge_length(s, nPP, pos, 0));
					t->camp[n][1][0][ui_pos] *= s->iv[IDX(s, nP, pos[0], pos[1], pos[2])];
					t->camp[n][1][1][ui_pos] *= s->iv[IDX(s, nPP, pos[0], pos[1], pos[2])];
				}
				if (t->active[n][0]) t->camp[n][0][0][ui_pos] *= -1;
				if (t->active[n][1]) t->camp[n][1][1][ui_pos] *= -1;
				if (pos[nP] == t->stop[nP]) {
					if (t->active[n][0]) t->camp[n][0][1][ui_pos] = 0;
					if (t->active[n][1]) t->camp[n][1][1][ui_pos] = 0;
				}
				if (pos[nPP] == t->stop[nPP]) {
					if (t->active[n][0]) t->camp[n][0][0][ui_pos] = 0;
					if (t->active[n][1]) t->camp[n][1][0][ui_pos] = 0;
				}
				/* voltage updates :307-365 */
				pos[n] = t->start[n] - 1;
				if (t->active[n][0]) {
					yee_coords(s, nP, pos, coord, 1); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(vdelay, vdd, vamp, 0, 1, 1.0, t->h_amp[nP] * orc_edge_length(s, nP, pos, 1));
					yee_coords(s, nPP, pos, coord, 1); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(vdelay, vdd, vamp, 0, 0, 1.0, t->h_amp[nPP] * orc_edge_length(s, nPP, pos, 1));
					++pos[n];
					t->vamp[n][0][0][ui_pos] *= s->vi[IDX(s, nP, pos[0], pos[1], pos[2])];
					t->vamp[n][0][1][ui_pos] *= s->vi[IDX(s, nPP, pos[0], pos[1], pos[2])];
				}
				pos[n] = t->stop[n];
				if (t->active[n][1]) {
					yee_coords(s, nP, pos, coord, 1); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(vdelay, vdd, vamp, 1, 1, 1.0, t->h_amp[nP] * orc_edge_length(s, nP, pos, 1));
					yee_coords(s, nPP, pos, coord, 1); dist = PW_DIST(coord, origin, t->prop_dir);
					TF_SET(vdelay, vdd, vamp, 1, 0, 1.0, t->h_amp[nPP] * orc_edge_length(s, nPP, pos, 1));
					t->vamp[n][1][0][ui_pos] *= s->vi[IDX(s, nP, pos[0], pos[1], pos[2])];
					t->vamp[n][1][1][ui_pos] *= s->vi[IDX(s, nPP, pos[0], pos[1], pos[2])];
				}
				if (t->active[n][1]) t->vamp[n][1][0][ui_pos] *= -1;
				if (t->active[n][0]) t->vamp[n][0][1][ui_pos] *= -1;
				if (pos[nP] == t->stop[nP]) {
					if (t->active[n][0]) t->vamp[n][0][0][ui_pos] = 0;
					if (t->active[n][1]) t->vamp[n][1][0][ui_pos] = 0;
				}
				if (pos[nPP] == t->stop[nPP]) {
					if (t->active[n][0]) t->vamp[n][0][1][ui_pos] = 0;
					if (t->active[n][1]) t->vamp[n][1][1][ui_pos] = 0;
				}
				++pos[nPP];
				++ui_pos;
			}
			++pos[nP];
		}
	}
	t->max_delay = max_delay + 1;
	t->lookup = xcalloc(t->max_delay + 1, sizeof(unsigned)); /* engine_ext_tfsf.cpp:25-27 */
	return 1;
}
/* the delay lookup of DoPostVoltageUpdates / DoPostCurrentUpdates, engine_ext_tfsf.cpp:38-53 / :128-141
   (the current version stops one entry earlier; that entry still holds the value the voltage hook of the
   same timestep wrote, which is the same expression) */
static void tfsf_lookup(orc_sim* s, tfsf_t* t, unsigned upto_incl)
{
	unsigned numTS = s->numTS, length = s->sig_len;
	int p = s->exc_period > 0 ? (int)(s->exc_period / s->dT) : 0;
	for (unsigned n = 0; n <= upto_incl; ++n) {
		if (numTS < n) t->lookup[n] = 0;
		else if (numTS - n >= length && p == 0) t->lookup[n] = 0;
		else t->lookup[n] = numTS - n;
		if (p > 0) t->lookup[n] = t->lookup[n] % p;
	}
}
static void tfsf_postV(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	if (tid != 0) return;
	tfsf_t* t = e->data;
	tfsf_lookup(s, t, t->max_delay);
	const float* signal = s->sig_i; /* "get the current signal since an H-field is added" */
	unsigned pos[3];
	for (int n = 0; n < 3; ++n) {
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		for (int l = 0; l < 2; ++l) {
			if (!t->active[n][l]) continue;
			pos[nP] = t->start[nP];
			unsigned u = 0;
			for (unsigned i = 0; i < t->nl[nP]; ++i) {
				pos[nPP] = t->start[nPP];
				for (unsigned j = 0; j < t->nl[nPP]; ++j) {
					pos[n] = l ? t->stop[n] : t->start[n];
					VOLT(s, nP, pos) = (float)(VOLT(s, nP, pos)
					    + (1.0 - t->vdd[n][l][0][u]) * t->vamp[n][l][0][u] * signal[t->lookup[t->vdelay[n][l][0][u]]]
					    + t->vdd[n][l][0][u] * t->vamp[n][l][0][u] * signal[t->lookup[1 + t->vdelay[n][l][0][u]]]);
					VOLT(s, nPP, pos) = (float)(VOLT(s, nPP, pos)
					    + (1.0 - t->vdd[n][l][1][u]) * t->vamp[n][l][1][u] * signal[t->lookup[t->vdelay[n][l][1][u]]]
					    + t->vdd[n][l][1][u] * t->vamp[n][l][1][u] * signal[t->lookup[1 + t->vdelay[n][l][1][u]]]);
					++pos[nPP];
					++u;
				}
				++pos[nP];
			}
		}
	}
}
static void tfsf_postI(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	if (tid != 0) return;
	tfsf_t* t = e->data;
	if (t->max_delay > 0) tfsf_lookup(s, t, t->max_delay - 1);
	const float* signal = s->sig_v;
	unsigned pos[3];
	for (int n = 0; n < 3; ++n) {
		if (!t->active[n][0] && !t->active[n][1]) continue;
		int nP = (n + 1) % 3, nPP = (n + 2) % 3;
		for (int l = 0; l < 2; ++l) {
			if (!t->active[n][l]) continue;
			pos[nP] = t->start[nP];
			unsigned u = 0;
			for (unsigned i = 0; i < t->nl[nP]; ++i) {
				pos[nPP] = t->start[nPP];
				for (unsigned j = 0; j < t->nl[nPP]; ++j) {
					pos[n] = l ? t->stop[n] : t->start[n] - 1;
					CURR(s, nP, pos) = (float)(CURR(s, nP, pos)
					    + (1.0 - t->cdd[n][l][0][u]) * t->camp[n][l][0][u] * signal[t->lookup[t->cdelay[n][l][0][u]]]
					    + t->cdd[n][l][0][u] * t->camp[n][l][0][u] * signal[t->lookup[1 + t->cdelay[n][l][0][u]]]);
					CURR(s, nPP, pos) = (float)(CURR(s, nPP, pos)
					    + (1.0 - t->cdd[n][l][1][u]) * t->camp[n][l][1][u] * signal[t->lookup[t->cdelay[n][l][1][u]]]
					    + t->cdd[n][l][1][u] * t->camp[n][l][1][u] * signal[t->lookup[1 + t->cdelay[n][l][1][u]]]);
					++pos[nPP];
					++u;
				}
				++pos[nP];
			}
		}
	}
}

/* ---- local absorbing sheets --------------------------------------------------------------
   openEMS::SetupAbsorbingSheets openems.cpp:411-441 (one extension per primitive, added after
   the lumped RLC extension, :1242-1243), Operator_Ext_Absorbing_BC::SetInitParams / BuildExtension
   operator_ext_absorbing_bc.cpp:60-245, Engine_Ext_Absorbing_BC engine_ext_absorbing_bc.cpp:38-366.
   x0/x1 are the snapped mesh indices of the sheet (m_sheetX0/X1). */
int orc_add_absorbing_sheet(orc_sim* s, const unsigned x0[3], const unsigned x1[3], int normal_positive, int type,
                            double phase_velocity)
{
	if (s->built || s->nabc >= 8) return -1;
	abc_t* a = &s->abc[s->nabc];
	memset(a, 0, sizeof(*a));
	int sheet = 0;
	a->ny = -1;
	for (int d = 0; d < 3; ++d) {
		a->x0[d] = x0[d]; a->x1[d] = x1[d];
		unsigned nc = x1[d] - x0[d] + 1;
		sheet += nc == 1;
		if (nc == 1) a->ny = d;
	}
	if (sheet != 1) return -2; /* "Absorbing sheet is not a sheet! Skipping." :112-117 */
	a->phase_velocity = phase_velocity == 0.0 ? C0 : phase_velocity; /* :120-127 */
	a->type = type; a->positive = normal_positive != 0;
	++s->nabc;
	return 0;
}
static void build_abc(orc_sim* s, abc_t* a)
{
	/* BuildExtension :138-245 */
	a->nyP = (a->ny + 1) % 3; a->nyPP = (a->ny + 2) % 3;
	unsigned pos[3] = {0, 0, 0};
	pos[a->ny] = a->x0[a->ny];
	const double delta = fabs(orc_edge_length(s, a->ny, pos, 0));
	const float vt = (float)(a->phase_velocity * s->dT); /* FDTD_FLOAT vt_nyP = m_phaseVelocity*dT */
	a->nl[0] = a->x1[a->nyP] - a->x0[a->nyP] + 1;
	a->nl[1] = a->x1[a->nyPP] - a->x0[a->nyPP] + 1;
	const size_t n = (size_t)a->nl[0] * a->nl[1];
	a->K1P = xcalloc(n, sizeof(float)); a->K1PP = xcalloc(n, sizeof(float));
	a->K2P = xcalloc(n, sizeof(float)); a->K2PP = xcalloc(n, sizeof(float));
	a->vP = xcalloc(n, sizeof(float)); a->vPP = xcalloc(n, sizeof(float));
	a->iP = xcalloc(n, sizeof(float)); a->iPP = xcalloc(n, sizeof(float));
	for (size_t q = 0; q < n; ++q) {
		a->K1P[q] = (float)((vt - delta) / (vt + delta));
		a->K1PP[q] = (float)((vt - delta) / (vt + delta));
		if (a->type == 2) { a->K2P[q] = (float)(vt / delta); a->K2PP[q] = (float)(vt / delta); }
	}
	/* Engine_Ext_Absorbing_BC ctor :64-71 */
	a->shift_V = a->x0[a->ny] + (a->positive ? 1 : -1);
	a->pos_I = a->x0[a->ny] + (a->positive ? 0 : -1);
	a->shift_I = a->x0[a->ny] + (a->positive ? 1 : -2);
}
static void abc_preV(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	if (tid != 0) return; /* SetNumberOfThreads(1) :79 */
	abc_t* a = e->data;
	unsigned pos[3] = {0, 0, 0}, ps[3] = {0, 0, 0};
	pos[a->ny] = a->x0[a->ny]; ps[a->ny] = a->shift_V;
	for (unsigned i = 0; i < a->nl[0]; ++i) {
		ps[a->nyP] = pos[a->nyP] = a->x0[a->nyP] + i;
		for (unsigned j = 0; j < a->nl[1]; ++j) {
			ps[a->nyPP] = pos[a->nyPP] = a->x0[a->nyPP] + j;
			size_t o = (size_t)i * a->nl[1] + j;
			a->vP[o] = VOLT(s, a->nyP, ps) - a->K1P[o] * VOLT(s, a->nyP, pos);
			a->vPP[o] = VOLT(s, a->nyPP, ps) - a->K1PP[o] * VOLT(s, a->nyPP, pos);
		}
	}
}
static void abc_postV(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	if (tid != 0) return;
	abc_t* a = e->data;
	unsigned ps[3] = {0, 0, 0};
	ps[a->ny] = a->shift_V;
	for (unsigned i = 0; i < a->nl[0]; ++i) {
		ps[a->nyP] = a->x0[a->nyP] + i;
		for (unsigned j = 0; j < a->nl[1]; ++j) {
			ps[a->nyPP] = a->x0[a->nyPP] + j;
			size_t o = (size_t)i * a->nl[1] + j;
			a->vP[o] += a->K1P[o] * VOLT(s, a->nyP, ps);
			a->vPP[o] += a->K1PP[o] * VOLT(s, a->nyPP, ps);
		}
	}
}
static void abc_applyV(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	if (tid != 0) return;
	abc_t* a = e->data;
	unsigned pos[3] = {0, 0, 0};
	pos[a->ny] = a->x0[a->ny];
	for (unsigned i = 0; i < a->nl[0]; ++i) {
		pos[a->nyP] = a->x0[a->nyP] + i;
		for (unsigned j = 0; j < a->nl[1]; ++j) {
			pos[a->nyPP] = a->x0[a->nyPP] + j;
			size_t o = (size_t)i * a->nl[1] + j;
			VOLT(s, a->nyP, pos) = a->vP[o];
			VOLT(s, a->nyPP, pos) = a->vPP[o];
		}
	}
}
static void abc_preI(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	abc_t* a = e->data;
	if (tid != 0 || a->type != 2) return; /* super-absorption only :232-233 */
	unsigned pos[3] = {0, 0, 0}, ps[3] = {0, 0, 0};
	pos[a->ny] = a->pos_I; ps[a->ny] = a->shift_I;
	for (unsigned i = 0; i + 1 < a->nl[0]; ++i) {
		ps[a->nyP] = pos[a->nyP] = a->x0[a->nyP] + i;
		for (unsigned j = 0; j + 1 < a->nl[1]; ++j) {
			ps[a->nyPP] = pos[a->nyPP] = a->x0[a->nyPP] + j;
			size_t o = (size_t)i * a->nl[1] + j;
			a->iP[o] = CURR(s, a->nyP, ps) - a->K1P[o] * CURR(s, a->nyP, pos);
			a->iPP[o] = CURR(s, a->nyPP, ps) - a->K1PP[o] * CURR(s, a->nyPP, pos);
		}
	}
}
static void abc_postI(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	abc_t* a = e->data;
	if (tid != 0 || a->type != 2) return;
	unsigned ps[3] = {0, 0, 0};
	ps[a->ny] = a->shift_I;
	for (unsigned i = 0; i + 1 < a->nl[0]; ++i) {
		ps[a->nyP] = a->x0[a->nyP] + i;
		for (unsigned j = 0; j + 1 < a->nl[1]; ++j) {
			ps[a->nyPP] = a->x0[a->nyPP] + j;
			size_t o = (size_t)i * a->nl[1] + j;
			a->iP[o] += a->K1P[o] * CURR(s, a->nyP, ps);
			a->iPP[o] += a->K1PP[o] * CURR(s, a->nyPP, ps);
		}
	}
}
static void abc_applyI(orc_sim* s, ext_t* e, int tid, int nth)
{
	(void)nth;
	abc_t* a = e->data;
	if (tid != 0 || a->type != 2) return;
	unsigned pos[3] = {0, 0, 0};
	pos[a->ny] = a->pos_I;
	for (unsigned i = 0; i + 1 < a->nl[0]; ++i) {
		pos[a->nyP] = a->x0[a->nyP] + i;
		for (unsigned j = 0; j + 1 < a->nl[1]; ++j) {
			pos[a->nyPP] = a->x0[a->nyPP] + j;
			size_t o = (size_t)i * a->nl[1] + j;
			/* (Hsa*K2 + Hc)/(K2 + 1.0): float numerator, double denominator :355-356 */
			CURR(s, a->nyP, pos) = (float)((a->iP[o] * a->K2P[o] + CURR(s, a->nyP, pos)) / (a->K2P[o] + 1.0));
			CURR(s, a->nyPP, pos) = (float)((a->iPP[o] * a->K2PP[o] + CURR(s, a->nyPP, pos)) / (a->K2PP[o] + 1.0));
		}
	}
}

/* Engine::SortExtensionByPriority engine.cpp:87-98: stable ascending sort, then reverse */
static void sort_exts(orc_sim* s)
{
	for (int i = 1; i < s->nexts; ++i) { /* insertion sort == stable */
		ext_t t = s->exts[i];
		int j = i - 1;
		while (j >= 0 && s->exts[j].prio > t.prio) { s->exts[j + 1] = s->exts[j]; --j; }
		s->exts[j + 1] = t;
	}
	for (int i = 0, j = s->nexts - 1; i < j; ++i, --j) {
		ext_t t = s->exts[i]; s->exts[i] = s->exts[j]; s->exts[j] = t;
	}
}

/* ---------------------------------------------------------------- build */
int orc_build(orc_sim* s, unsigned max_ts)
{
	if (s->built) return -1;
	size_t nc = ncell(s);
	s->EC_C = xcalloc(3 * nc, sizeof(float)); s->EC_G = xcalloc(3 * nc, sizeof(float));
	s->EC_L = xcalloc(3 * nc, sizeof(float)); s->EC_R = xcalloc(3 * nc, sizeof(float));
	/* Operator::Calc_EC_Range operator.cpp:1832-1860: EC arrays are FDTD_FLOAT */
#pragma omp parallel for collapse(2) schedule(static)
	for (unsigned i = 0; i < s->N[0]; ++i)
		for (unsigned j = 0; j < s->N[1]; ++j)
			for (unsigned k = 0; k < s->N[2]; ++k) {
				unsigned pos[3] = {i, j, k};
				size_t ip = ((size_t)i * s->N[1] + j) * s->N[2] + k;
				for (int n = 0; n < 3; ++n) {
					double EC[4];
					calc_ec_pos(s, n, pos, EC);
					s->EC_C[n * nc + ip] = EC[0]; s->EC_G[n * nc + ip] = EC[1];
					s->EC_L[n * nc + ip] = EC[2]; s->EC_R[n * nc + ip] = EC[3];
				}
			}
	/* timestep, operator.cpp:994-1025 */
	if (s->forced_dT > 0) s->dT = s->forced_dT;
	else s->dT = calc_timestep(s);
	s->dT *= s->ts_factor;
	if (s->exc_period > 0) {
		unsigned TS = (unsigned)ceil(s->exc_period / s->dT);
		s->dT = s->exc_period / TS;
	}
	s->vv = xcalloc(3 * nc, sizeof(float)); s->vi = xcalloc(3 * nc, sizeof(float));
	s->ii = xcalloc(3 * nc, sizeof(float)); s->iv = xcalloc(3 * nc, sizeof(float));
#pragma omp parallel for collapse(2) schedule(static)
	for (unsigned i = 0; i < s->N[0]; ++i)
		for (unsigned j = 0; j < s->N[1]; ++j)
			for (unsigned k = 0; k < s->N[2]; ++k) {
				unsigned pos[3] = {i, j, k};
				for (int n = 0; n < 3; ++n) calc_ecop_pos(s, n, pos);
			}
	int PEC[6], PMC[6];
	for (int n = 0; n < 6; ++n) { PEC[n] = s->bc[n] != -1; PMC[n] = s->bc[n] == 1; }
	apply_electric_bc(s, PEC);
	calc_pec(s);
	calc_lumped(s);
	apply_magnetic_bc(s, PMC);

	/* the excitation signal length is needed by the Mur start delay; the reference builds the
	   signal after CalcECOperator (openems.cpp:1271) but before CreateEngine (:1316) */
	if (build_signal(s, max_ts) != 0) return -2;

	/* extension builds in insertion order openems.cpp:1186-1243,388-406:
	   Excitation, Mur(xmin..zmax), UPML(xmin..zmax), Lorentz, LumpedRLC */
	s->nexts = 0;
	build_excitation(s);
	add_ext(s, PRIO_EXCITATION, NULL, NULL, NULL, exc_applyV, NULL, NULL, exc_applyI);
	/* Operator_Ext_TFSF is the second extension (openems.cpp:1187-1188); it stays inactive without a plane wave */
	if (s->tfsf.on && build_tfsf(s)) add_ext(s, PRIO_TFSF, &s->tfsf, NULL, tfsf_postV, NULL, NULL, tfsf_postI, NULL);
	s->nmur = 0;
	for (int n = 0; n < 6; ++n)
		if (s->bc[n] == 2) {
			mur_t* m = &s->mur[s->nmur++];
			build_mur(s, m, n / 2, n % 2);
			add_ext(s, PRIO_DEFAULT, m, mur_preV, mur_postV, mur_applyV, NULL, NULL, NULL);
		}
	create_upml(s);
	for (int b = 0; b < s->nupml; ++b) {
		build_upml_box(s, &s->upml[b]);
		add_ext(s, PRIO_UPML, &s->upml[b], upml_preV, upml_postV, NULL, upml_preI, upml_postI, NULL);
	}
	/* Operator_Ext_SteadyState is inserted after UPML and before Lorentz (openems.cpp:1206-1236) */
	if (s->ss) add_ext(s, PRIO_STEADYSTATE, s->ss, NULL, NULL, ss_applyV, NULL, NULL, NULL);
	build_lorentz(s);
	if (s->lor_order > 0)
		add_ext(s, PRIO_DEFAULT, NULL, lor_preV, NULL, lor_applyV, lor_preI, NULL, lor_applyI);
	for (int r = 0; r < s->nrlc; ++r)
		add_ext(s, PRIO_DEFAULT, &s->rlc[r], rlc_preV, NULL, rlc_applyV, NULL, NULL, NULL);
	/* absorbing sheets come last, openems.cpp:1242-1243 */
	for (int a = 0; a < s->nabc; ++a) {
		build_abc(s, &s->abc[a]);
		add_ext(s, PRIO_DEFAULT, &s->abc[a], abc_preV, abc_postV, abc_applyV, abc_preI, abc_postI, abc_applyI);
	}
	sort_exts(s);

	s->volt = xcalloc(3 * nc, sizeof(float));
	s->curr = xcalloc(3 * nc, sizeof(float));
	s->numTS = 0;
	s->built = 1;
	return 0;
}

/* ---------------------------------------------------------------- engine */
/* Engine::UpdateVoltages engine.cpp:110-168 */
static void update_voltages(orc_sim* s)
{
	const unsigned Nx = s->N[0], Ny = s->N[1], Nz = s->N[2];
	float *volt = s->volt, *curr = s->curr;
	const float *vv = s->vv, *vi = s->vi;
#define A(arr, n, i, j, k) arr[((((size_t)(n)) * Nx + (i)) * Ny + (j)) * Nz + (k)]
	for (unsigned i = 0; i < Nx; ++i) {
		unsigned s0 = i > 0;
		for (unsigned j = 0; j < Ny; ++j) {
			unsigned s1 = j > 0;
			for (unsigned k = 0; k < Nz; ++k) {
				unsigned s2 = k > 0;
				A(volt, 0, i, j, k) *= A(vv, 0, i, j, k);
				A(volt, 0, i, j, k) += A(vi, 0, i, j, k) * (A(curr, 2, i, j, k) - A(curr, 2, i, j - s1, k) - A(curr, 1, i, j, k) + A(curr, 1, i, j, k - s2));
				A(volt, 1, i, j, k) *= A(vv, 1, i, j, k);
				A(volt, 1, i, j, k) += A(vi, 1, i, j, k) * (A(curr, 0, i, j, k) - A(curr, 0, i, j, k - s2) - A(curr, 2, i, j, k) + A(curr, 2, i - s0, j, k));
				A(volt, 2, i, j, k) *= A(vv, 2, i, j, k);
				A(volt, 2, i, j, k) += A(vi, 2, i, j, k) * (A(curr, 1, i, j, k) - A(curr, 1, i - s0, j, k) - A(curr, 0, i, j, k) + A(curr, 0, i, j - s1, k));
			}
		}
	}
}
/* Engine::UpdateCurrents engine.cpp:170-222 */
static void update_currents(orc_sim* s)
{
	const unsigned Nx = s->N[0], Ny = s->N[1], Nz = s->N[2];
	float *volt = s->volt, *curr = s->curr;
	const float *ii = s->ii, *iv = s->iv;
	for (unsigned i = 0; i < Nx - 1; ++i)
		for (unsigned j = 0; j < Ny - 1; ++j)
			for (unsigned k = 0; k < Nz - 1; ++k) {
				A(curr, 0, i, j, k) *= A(ii, 0, i, j, k);
				A(curr, 0, i, j, k) += A(iv, 0, i, j, k) * (A(volt, 2, i, j, k) - A(volt, 2, i, j + 1, k) - A(volt, 1, i, j, k) + A(volt, 1, i, j, k + 1));
				A(curr, 1, i, j, k) *= A(ii, 1, i, j, k);
				A(curr, 1, i, j, k) += A(iv, 1, i, j, k) * (A(volt, 0, i, j, k) - A(volt, 0, i, j, k + 1) - A(volt, 2, i, j, k) + A(volt, 2, i + 1, j, k));
				A(curr, 2, i, j, k) *= A(ii, 2, i, j, k);
				A(curr, 2, i, j, k) += A(iv, 2, i, j, k) * (A(volt, 1, i, j, k) - A(volt, 1, i + 1, j, k) - A(volt, 0, i, j, k) + A(volt, 0, i, j + 1, k));
			}
#undef A
}

/* Engine::IterateTS engine.cpp:267-286, hook order engine.cpp:224-265 */
void orc_iterate(orc_sim* s, unsigned n_ts)
{
	unsigned old_csr = _mm_getcsr();
	_mm_setcsr(old_csr | 0x8040); /* tools/denormal.h:19-30 */
	for (unsigned it = 0; it < n_ts; ++it) {
		for (int n = s->nexts - 1; n >= 0; --n) if (s->exts[n].preV) s->exts[n].preV(s, &s->exts[n], 0, 1);
		update_voltages(s);
		for (int n = 0; n < s->nexts; ++n) if (s->exts[n].postV) s->exts[n].postV(s, &s->exts[n], 0, 1);
		for (int n = 0; n < s->nexts; ++n) if (s->exts[n].applyV) s->exts[n].applyV(s, &s->exts[n], 0, 1);
		for (int n = s->nexts - 1; n >= 0; --n) if (s->exts[n].preI) s->exts[n].preI(s, &s->exts[n], 0, 1);
		update_currents(s);
		for (int n = 0; n < s->nexts; ++n) if (s->exts[n].postI) s->exts[n].postI(s, &s->exts[n], 0, 1);
		for (int n = 0; n < s->nexts; ++n) if (s->exts[n].applyI) s->exts[n].applyI(s, &s->exts[n], 0, 1);
		++s->numTS;
	}
	_mm_setcsr(old_csr);
}

unsigned orc_num_ts(const orc_sim* s) { return s->numTS; }
float* orc_volt(orc_sim* s) { return s->volt; }
float* orc_curr(orc_sim* s) { return s->curr; }
const float* orc_upml_flux(const orc_sim* s, int b, int is_curr)
{
	return is_curr ? s->upml[b].curr_flux : s->upml[b].volt_flux;
}
void orc_reset_fields(orc_sim* s)
{
	size_t nc = ncell(s);
	memset(s->volt, 0, 3 * nc * sizeof(float));
	memset(s->curr, 0, 3 * nc * sizeof(float));
	s->numTS = 0;
	for (int b = 0; b < s->nupml; ++b) {
		size_t cnt = (size_t)3 * s->upml[b].n[0] * s->upml[b].n[1] * s->upml[b].n[2];
		memset(s->upml[b].volt_flux, 0, cnt * sizeof(float));
		memset(s->upml[b].curr_flux, 0, cnt * sizeof(float));
	}
	for (int m = 0; m < s->nmur; ++m) {
		size_t cnt = (size_t)s->mur[m].n[0] * s->mur[m].n[1];
		memset(s->mur[m].vP, 0, cnt * sizeof(float));
		memset(s->mur[m].vPP, 0, cnt * sizeof(float));
	}
	for (int o = 0; o < s->lor_order; ++o)
		for (int n = 0; n < 3; ++n) {
			lor_order_t* L = &s->lor[o];
			if (L->volt_ADE[n]) memset(L->volt_ADE[n], 0, L->count * sizeof(float));
			if (L->curr_ADE[n]) memset(L->curr_ADE[n], 0, L->count * sizeof(float));
			if (L->volt_Lor_ADE[n]) memset(L->volt_Lor_ADE[n], 0, L->count * sizeof(float));
			if (L->curr_Lor_ADE[n]) memset(L->curr_Lor_ADE[n], 0, L->count * sizeof(float));
		}
	for (int r = 0; r < s->nrlc; ++r) {
		rlc_t* R = &s->rlc[r];
		memset(R->Il, 0, R->count * sizeof(float));
		for (int n = 0; n < 3; ++n) {
			memset(R->Vdn[n], 0, R->count * sizeof(float));
			memset(R->Jn[n], 0, R->count * sizeof(float));
		}
	}
}

/* ---------------------------------------------------------------- accessors */
double orc_dT(const orc_sim* s) { return s->dT; }
unsigned orc_nyquist(const orc_sim* s) { return s->nyquist; }
const float* orc_coeff(const orc_sim* s, int w)
{
	return w == 0 ? s->vv : w == 1 ? s->vi : w == 2 ? s->ii : s->iv;
}
unsigned orc_signal_length(const orc_sim* s) { return s->sig_len; }
const float* orc_signal(const orc_sim* s, int is_curr) { return is_curr ? s->sig_i : s->sig_v; }
unsigned orc_signal_period_ts(const orc_sim* s)
{
	return s->exc_period > 0 ? (unsigned)(int)(s->exc_period / s->dT) : 0;
}
unsigned orc_exc_count(const orc_sim* s, int is_curr) { return is_curr ? s->ccount : s->vcount; }
void orc_exc_get(const orc_sim* s, int is_curr, unsigned* idx, unsigned* dir, float* amp, unsigned* delay)
{
	unsigned cnt = is_curr ? s->ccount : s->vcount;
	unsigned* const* I = is_curr ? s->cidx : s->vidx;
	for (int n = 0; n < 3; ++n) memcpy(idx + (size_t)n * cnt, I[n], cnt * sizeof(unsigned));
	memcpy(dir, is_curr ? s->cdir : s->vdir, cnt * sizeof(unsigned));
	memcpy(amp, is_curr ? s->camp : s->vamp, cnt * sizeof(float));
	memcpy(delay, is_curr ? s->cdelay : s->vdelay, cnt * sizeof(unsigned));
}
int orc_upml_count(const orc_sim* s) { return s->nupml; }
void orc_upml_box(const orc_sim* s, int b, unsigned start[3], unsigned nl[3])
{
	for (int n = 0; n < 3; ++n) { start[n] = s->upml[b].start[n]; nl[n] = s->upml[b].n[n]; }
}
const float* orc_upml_coeff(const orc_sim* s, int b, int w) { return s->upml[b].c[w]; }
int orc_mur_count(const orc_sim* s) { return s->nmur; }
void orc_mur_info(const orc_sim* s, int m, int* ny, int* top, unsigned* line, unsigned* shift,
                  unsigned nl[2], unsigned* start_ts)
{
	const mur_t* M = &s->mur[m];
	*ny = M->ny; *top = M->top; *line = M->line; *shift = M->shift;
	nl[0] = M->n[0]; nl[1] = M->n[1]; *start_ts = M->start_ts;
}
const float* orc_mur_coeff(const orc_sim* s, int m, int w) { return w ? s->mur[m].cPP : s->mur[m].cP; }
int orc_lorentz_order(const orc_sim* s) { return s->lor_order; }
unsigned orc_lorentz_count(const orc_sim* s, int o) { return s->lor[o].count; }
int orc_lorentz_flags(const orc_sim* s, int o)
{
	const lor_order_t* L = &s->lor[o];
	return (L->volt_on ? 1 : 0) | (L->curr_on ? 2 : 0) | (L->volt_lor_on ? 4 : 0) | (L->curr_lor_on ? 8 : 0);
}
const unsigned* orc_lorentz_pos(const orc_sim* s, int o, int n) { return s->lor[o].pos[n]; }
const float* orc_lorentz_coeff(const orc_sim* s, int o, int w, int n)
{
	const lor_order_t* L = &s->lor[o];
	switch (w) {
	case 0: return L->v_int[n];
	case 1: return L->v_ext[n];
	case 2: return L->v_lor[n];
	case 3: return L->i_int[n];
	case 4: return L->i_ext[n];
	default: return L->i_lor[n];
	}
}

/* ---------------------------------------------------------------- readout */
/* Engine_Interface_FDTD::CalcVoltageIntegral engine_interface_fdtd.cpp:206-232 */
double orc_voltage_integral(const orc_sim* s, const unsigned start[3], const unsigned stop[3])
{
	if (((start[0] != stop[0]) + (start[1] != stop[1]) + (start[2] != stop[2])) != 1) return 0;
	double result = 0;
	for (int n = 0; n < 3; ++n) {
		if (start[n] < stop[n]) {
			unsigned pos[3] = {start[0], start[1], start[2]};
			for (; pos[n] < stop[n]; ++pos[n]) result += VOLT(s, n, pos);
		} else {
			unsigned pos[3] = {stop[0], stop[1], stop[2]};
			for (; pos[n] < start[n]; ++pos[n]) result -= VOLT(s, n, pos);
		}
	}
	return result;
}

/* ProcessCurrent::CalcIntegral Common/processcurrent.cpp:96-171 (fp32 accumulator) */
double orc_current_integral(const orc_sim* s, const unsigned start[3], const unsigned stop[3],
                            int nd, const int si[3], const int ei[3])
{
	float current = 0;
#define GC(n, a, b, c) s->curr[IDX(s, n, a, b, c)]
	switch (nd) {
	case 0:
		if (ei[0] && si[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) current += GC(1, stop[0], i, start[2]);
		if (ei[0] && ei[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) current += GC(2, stop[0], stop[1], i);
		if (si[0] && ei[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) current -= GC(1, start[0], i, stop[2]);
		if (si[0] && si[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) current -= GC(2, start[0], start[1], i);
		break;
	case 1:
		if (si[0] && si[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) current += GC(2, start[0], start[1], i);
		if (ei[1] && ei[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) current += GC(0, i, stop[1], stop[2]);
		if (ei[0] && ei[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) current -= GC(2, stop[0], stop[1], i);
		if (si[1] && si[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) current -= GC(0, i, start[1], start[2]);
		break;
	case 2:
		if (si[1] && si[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) current += GC(0, i, start[1], start[2]);
		if (ei[0] && si[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) current += GC(1, stop[0], i, start[2]);
		if (ei[1] && ei[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) current -= GC(0, i, stop[1], stop[2]);
		if (si[0] && ei[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) current -= GC(1, start[0], i, stop[2]);
		break;
	default:
		return 0.0;
	}
#undef GC
	return current;
}

/* GetRawField type 0 engine_interface_fdtd.cpp:263-268 ; GetRawDualField type 0 :136-141 */
static double raw_field(const orc_sim* s, int is_H, int n, const unsigned pos[3])
{
	double value = is_H ? CURR(s, n, pos) : VOLT(s, n, pos);
	double delta = orc_edge_length(s, n, pos, is_H);
	if (delta) return value / delta;
	return 0.0;
}
void orc_raw_field(const orc_sim* s, int is_H, const unsigned pos[3], double out[3])
{
	for (int n = 0; n < 3; ++n) out[n] = raw_field(s, is_H, n, pos);
}

/* Engine_Interface_FDTD::CalcFastEnergy (scalar engine branch) engine_interface_fdtd.cpp:302-347 */
double orc_energy(const orc_sim* s)
{
	double E = 0, H = 0;
	unsigned pos[3];
	for (pos[0] = 0; pos[0] < s->N[0] - 1; ++pos[0])
		for (pos[1] = 0; pos[1] < s->N[1] - 1; ++pos[1])
			for (pos[2] = 0; pos[2] < s->N[2] - 1; ++pos[2])
				for (int n = 0; n < 3; ++n) {
					E += VOLT(s, n, pos) * VOLT(s, n, pos);
					H += CURR(s, n, pos) * CURR(s, n, pos);
				}
	return EPS0 * E + MUE0 * H;
}

/* GetRawInterpolatedField engine_interface_fdtd.cpp:63-124 */
static void interp_E(const orc_sim* s, int interp, const unsigned pos[3], double out[3])
{
	unsigned ip[3] = {pos[0], pos[1], pos[2]};
	switch (interp) {
	default:
	case 0:
		for (int n = 0; n < 3; ++n) out[n] = raw_field(s, 0, n, pos);
		break;
	case 1:
		for (int n = 0; n < 3; ++n) {
			if (pos[n] == s->N[n] - 1) {
				--ip[n]; out[n] = raw_field(s, 0, n, ip); ++ip[n];
				continue;
			}
			double delta = orc_edge_length(s, n, ip, 0);
			out[n] = raw_field(s, 0, n, ip);
			if (delta == 0) { out[n] = 0; continue; }
			if (pos[n] == 0) continue;
			--ip[n];
			double dDown = orc_edge_length(s, n, ip, 0);
			double dRel = delta / (delta + dDown);
			out[n] = out[n] * (1.0 - dRel) + raw_field(s, 0, n, ip) * dRel;
			++ip[n];
		}
		break;
	case 2:
		for (int n = 0; n < 3; ++n) {
			int nP = (n + 1) % 3, nPP = (n + 2) % 3;
			if (pos[0] == s->N[0] - 1 || pos[1] == s->N[1] - 1 || pos[2] == s->N[2] - 1) { out[n] = 0; continue; }
			out[n] = raw_field(s, 0, n, ip);
			++ip[nP]; out[n] += raw_field(s, 0, n, ip);
			++ip[nPP]; out[n] += raw_field(s, 0, n, ip);
			--ip[nP]; out[n] += raw_field(s, 0, n, ip);
			--ip[nPP];
			out[n] /= 4;
		}
		break;
	}
}
/* GetRawInterpolatedDualField engine_interface_fdtd.cpp:150-204 */
static void interp_H(const orc_sim* s, int interp, const unsigned pos[3], double out[3])
{
	unsigned ip[3] = {pos[0], pos[1], pos[2]};
	switch (interp) {
	default:
	case 0:
		for (int n = 0; n < 3; ++n) out[n] = raw_field(s, 1, n, pos);
		break;
	case 1:
		for (int n = 0; n < 3; ++n) {
			int nP = (n + 1) % 3, nPP = (n + 2) % 3;
			if (pos[0] == s->N[0] - 1 || pos[1] == s->N[1] - 1 || pos[2] == s->N[2] - 1 || pos[nP] == 0 || pos[nPP] == 0) { out[n] = 0; continue; }
			out[n] = raw_field(s, 1, n, ip);
			--ip[nP]; out[n] += raw_field(s, 1, n, ip);
			--ip[nPP]; out[n] += raw_field(s, 1, n, ip);
			++ip[nP]; out[n] += raw_field(s, 1, n, ip);
			++ip[nPP];
			out[n] /= 4;
		}
		break;
	case 2:
		for (int n = 0; n < 3; ++n) {
			double delta = orc_edge_length(s, n, ip, 1);
			out[n] = raw_field(s, 1, n, ip);
			if (pos[n] >= s->N[n] - 1) { out[n] = 0; continue; }
			++ip[n];
			double dUp = orc_edge_length(s, n, ip, 1);
			double dRel = delta / (delta + dUp);
			out[n] = out[n] * (1.0 - dRel) + raw_field(s, 1, n, ip) * dRel;
			--ip[n];
		}
		break;
	}
}

/* ProcessFields::CalcField Common/processfields.cpp:283-409 over the full index range
   start..stop (no sub-sampling), stored x-fastest like tools/hdf5_file_writer.cpp:286-302 */
void orc_dump_field(const orc_sim* s, int is_H, int interp, const unsigned start[3],
                    const unsigned stop[3], float* out)
{
	unsigned n[3] = {stop[0] - start[0] + 1, stop[1] - start[1] + 1, stop[2] - start[2] + 1};
	size_t cnt = (size_t)n[0] * n[1] * n[2];
	for (unsigned k = 0; k < n[2]; ++k)
		for (unsigned j = 0; j < n[1]; ++j)
			for (unsigned i = 0; i < n[0]; ++i) {
				unsigned pos[3] = {start[0] + i, start[1] + j, start[2] + k};
				double o[3];
				if (is_H) interp_H(s, interp, pos, o); else interp_E(s, interp, pos, o);
				size_t off = ((size_t)k * n[1] + j) * n[0] + i;
				out[off] = (float)o[0]; out[cnt + off] = (float)o[1]; out[2 * cnt + off] = (float)o[2];
			}
}


/* ProcessFieldsFD::Process Common/processfields_fd.cpp:72-107 -- the running DFT of a field dump.
   weight: exp_jwt_2_dt of lines 84-86: std::exp((complex<float>)(-2.0*_I*M_PI*f*T)), then *= 2
   (single-sided spectrum), then *= Op->GetTimestep()*m_FD_Interval (converted to float by
   complex<float>::operator*=).  std::exp(complex<float>) is cexpf. */
#include <complex.h>
void orc_fd_weight(double freq, double T, double dT, unsigned interval, float out[2])
{
	const double complex arg = -2.0 * I * M_PI * freq * T;
	float complex e = cexpf((float complex)arg);
	float re = crealf(e), im = cimagf(e);
	re *= 2.0f; im *= 2.0f;
	const float sc = (float)(dT * (double)interval);
	re *= sc; im *= sc;
	out[0] = re; out[1] = im;
}
/* lines 88-100: field_fd += field_td * exp_jwt_2_dt, complex<float> += float * complex<float>
   (operator*(const float&, const complex<float>&) scales both parts); acc is interleaved re/im */
void orc_fd_accumulate(float* acc, const float* td, size_t n, const float w[2])
{
	for (size_t t = 0; t < n; ++t) {
		const float pr = td[t] * w[0], pi = td[t] * w[1];
		acc[2 * t] = acc[2 * t] + pr;
		acc[2 * t + 1] = acc[2 * t + 1] + pi;
	}
}


/* ProcessModeMatch::CalcMultipleIntegrals Common/processmodematch.cpp:222-266 on the surface
   start..stop (already sorted / pulled off the boundaries as InitProcess does, lines 86-104) with
   the normalised mode template dist0/dist1 [posP][posPP]; node interpolation (line 82);
   area = Operator::GetNodeArea operator.cpp:235-240 = GetNodeWidth(nP)*GetNodeWidth(nPP),
   GetNodeWidth = GetEdgeLength(ny,pos,!dualMesh) operator.h:174 */
void orc_mode_match(const orc_sim* s, int is_H, int ny, const unsigned start[3], const unsigned stop[3],
                    const double* dist0, const double* dist1, double out2[2])
{
	const int nP = (ny + 1) % 3, nPP = (ny + 2) % 3;
	const unsigned nl0 = stop[nP] - start[nP] + 1, nl1 = stop[nPP] - start[nPP] + 1;
	double value = 0, purity = 0;
	unsigned pos[3] = {0, 0, 0};
	pos[ny] = start[ny];
	for (unsigned posP = 0; posP < nl0; ++posP) {
		pos[nP] = start[nP] + posP;
		for (unsigned posPP = 0; posPP < nl1; ++posPP) {
			pos[nPP] = start[nPP] + posPP;
			const double area = orc_edge_length(s, nP, pos, !is_H) * orc_edge_length(s, nPP, pos, !is_H);
			double o[3];
			if (is_H) interp_H(s, 1, pos, o); else interp_E(s, 1, pos, o);
			const double* dist[2] = {dist0, dist1};
			for (int n = 0; n < 2; ++n) {
				const double field = o[(ny + n + 1) % 3];
				value += field * dist[n][(size_t)posP * nl1 + posPP] * area;
				purity += field * field * area;
			}
		}
	}
	out2[1] = purity != 0 ? value * value / purity : 0;
	out2[0] = value;
}


/* absorbing sheet tables for the engine upload */
int orc_abc_count(const orc_sim* s) { return s->nabc; }
void orc_abc_info(const orc_sim* s, int a, int* ny, int* type, int* positive, unsigned x0[3], unsigned x1[3])
{
	const abc_t* A = &s->abc[a];
	*ny = A->ny; *type = A->type; *positive = A->positive;
	for (int d = 0; d < 3; ++d) { x0[d] = A->x0[d]; x1[d] = A->x1[d]; }
}
void orc_abc_coeff(const orc_sim* s, int a, float* K1P, float* K1PP, float* K2P, float* K2PP)
{
	const abc_t* A = &s->abc[a];
	const size_t n = (size_t)A->nl[0] * A->nl[1];
	memcpy(K1P, A->K1P, n * sizeof(float)); memcpy(K1PP, A->K1PP, n * sizeof(float));
	memcpy(K2P, A->K2P, n * sizeof(float)); memcpy(K2PP, A->K2PP, n * sizeof(float));
}


/* TFSF tables for the engine upload: which = 0 voltage, 1 current; returns the point count of face (n, l) */
int orc_tfsf_on(const orc_sim* s) { return s->tfsf.on && s->tfsf.lookup != NULL; }
unsigned orc_tfsf_max_delay(const orc_sim* s) { return s->tfsf.max_delay; }
void orc_tfsf_box(const orc_sim* s, unsigned start[3], unsigned stop[3], int active[6])
{
	for (int n = 0; n < 3; ++n) { start[n] = s->tfsf.start[n]; stop[n] = s->tfsf.stop[n]; active[2 * n] = s->tfsf.active[n][0]; active[2 * n + 1] = s->tfsf.active[n][1]; }
}
unsigned orc_tfsf_face(const orc_sim* s, int which, int n, int l, int c, unsigned* delay, float* delta, float* amp)
{
	const tfsf_t* t = &s->tfsf;
	if (!t->active[n][l]) return 0;
	const unsigned numP = t->nl[(n + 1) % 3] * t->nl[(n + 2) % 3];
	memcpy(delay, which ? t->cdelay[n][l][c] : t->vdelay[n][l][c], numP * sizeof(unsigned));
	memcpy(delta, which ? t->cdd[n][l][c] : t->vdd[n][l][c], numP * sizeof(float));
	memcpy(amp, which ? t->camp[n][l][c] : t->vamp[n][l][c], numP * sizeof(float));
	return numP;
}
