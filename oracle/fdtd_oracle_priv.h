/* fdtd_oracle_priv.h -- private structures shared by fdtd_oracle.c and fdtd_oracle_sse.c.
 * TEST INFRASTRUCTURE ONLY (see fdtd_oracle.h). */
#ifndef FDTD_ORACLE_PRIV_H
#define FDTD_ORACLE_PRIV_H
#include "fdtd_oracle.h"
#include <stddef.h>

#define EPS0 8.85418781762e-12 /* tools/constants.h:23 */
#define MUE0 1.256637062e-6    /* tools/constants.h:24 */
#define C0 299792458.0         /* tools/constants.h:25 */
#define Z0 376.730313461       /* tools/constants.h:26 */
#define ORC_PI 3.141592653589793238462643383279

#define MAX_ORDER 8

enum { P_MATERIAL = 0, P_METAL, P_LORENTZ, P_EXCITATION };

typedef struct {
	int type, prio;
	double start[3], stop[3];
	double epsR, mueR, kappa, sigma;
	int order;
	double eps_fp[MAX_ORDER], eps_tau[MAX_ORDER], eps_flor[MAX_ORDER];
	double mue_fp[MAX_ORDER], mue_tau[MAX_ORDER], mue_flor[MAX_ORDER];
	int exc_type;
	double exc_vec[3], delay;
} prop_t;

typedef struct {
	double start[3], stop[3];
	int dir, caps;
	double R, C;
} lumped_t;

typedef struct {
	unsigned start[3], n[3];
	float *c[6]; /* vv vvfn vvfo ii iifn iifo, each [3][nx][ny][nz] local NIJK */
	float *volt_flux, *curr_flux;
} upml_t;

typedef struct {
	int ny, nyP, nyPP, top;
	unsigned line, shift, n[2], start_ts;
	float *cP, *cPP;       /* ArrayIJ [i][j], j fastest */
	float *vP, *vPP;       /* engine state */
} mur_t;

/* Operator_Ext_TFSF / Engine_Ext_TFSF: total-field / scattered-field plane wave box */
typedef struct {
	int on;
	unsigned start[3], stop[3], nl[3];
	double prop_dir[3], e_amp[3], h_amp[3], ph_vel;
	int active[3][2];
	unsigned max_delay;
	unsigned* vdelay[3][2][2]; float* vdd[3][2][2]; float* vamp[3][2][2];
	unsigned* cdelay[3][2][2]; float* cdd[3][2][2]; float* camp[3][2][2];
	unsigned* lookup;
} tfsf_t;

/* Operator_Ext_Absorbing_BC / Engine_Ext_Absorbing_BC: local absorbing sheet */
typedef struct {
	int ny, nyP, nyPP, type, positive;     /* type 1: MUR_1ST, 2: MUR_1ST_SA */
	unsigned x0[3], x1[3], nl[2];
	double phase_velocity;
	unsigned shift_V, pos_I, shift_I;
	float *K1P, *K1PP, *K2P, *K2PP;        /* ArrayIJ [i][j] */
	float *vP, *vPP, *iP, *iPP;            /* engine state */
} abc_t;

typedef struct {
	unsigned count;
	int volt_on, curr_on, volt_lor_on, curr_lor_on;
	unsigned* pos[3];
	float *v_int[3], *v_ext[3], *v_lor[3], *i_int[3], *i_ext[3], *i_lor[3];
	float *volt_ADE[3], *curr_ADE[3], *volt_Lor_ADE[3], *curr_Lor_ADE[3];
} lor_order_t;

typedef struct {
	unsigned count;
	int* dir;
	unsigned* pos[3];
	float *ilv, *i2v, *vvd, *vv2, *vj1, *vj2, *ib0, *b1, *b2;
	float *Vdn[3], *Jn[3], *Il;
} rlc_t;

typedef struct {
	unsigned period, count;
	int* dir;
	unsigned* pos[3];
	double* rec;          /* [count][2*period] */
	double last_max_diff, last_total_energy;
} ss_t;

struct ext_s;
typedef void (*hook_fn)(orc_sim*, struct ext_s*, int tid, int nth);
typedef struct ext_s {
	int prio;
	void* data;
	hook_fn preV, postV, applyV, preI, postI, applyI;
} ext_t;

/* FDTD/extensions/engine_extension.h:21-29 */
#define PRIO_DEFAULT 0
#define PRIO_UPML 1000000
#define PRIO_TFSF 50000
#define PRIO_EXCITATION (-1000)
#define PRIO_STEADYSTATE 2000000

struct orc_sim {
	unsigned N[3];
	double* lines[3];
	double grid_delta;
	int bc[6];
	unsigned pml_size[6];
	double bg[4];
	double mur_vphase;
	double forced_dT, ts_factor;

	prop_t* props; int nprops;
	lumped_t* lumped; int nlumped;

	/* operator */
	double dT;
	float *EC_C, *EC_G, *EC_L, *EC_R;
	float *vv, *vi, *ii, *iv;
	/* excitation signal */
	int exc_kind; double exc_f0, exc_fc, exc_fmax, exc_period;
	unsigned sig_len, nyquist;
	float *sig_v, *sig_i;
	/* excitation lists */
	unsigned vcount, ccount;
	unsigned *vidx[3], *vdir, *vdelay; float* vamp;
	unsigned *cidx[3], *cdir, *cdelay; float* camp;

	upml_t upml[6]; int nupml;
	mur_t mur[6]; int nmur;
	int lor_order; lor_order_t lor[MAX_ORDER];
	rlc_t* rlc; int nrlc;
	ss_t* ss;
	abc_t abc[8]; int nabc;
	tfsf_t tfsf;

	/* engine */
	float *volt, *curr;
	/* sse-compressed engine state (fdtd_oracle_sse.c); sse!=0 switches the field accessors */
	int sse; unsigned nv; float *f4_volt, *f4_curr;
	unsigned numTS;
	ext_t exts[32]; int nexts;
	int built;
};


/* field accessors: ArrayNIJK for the scalar engine (FDTD/engine.h:55-101), ArrayENG
   I-J-K-N with 4 interleaved z lanes for the sse engines (FDTD/engine_sse.h:38-84,
   tools/arraylib/array_e.h:57-60) */
static inline size_t orc_idx(const orc_sim* s, int n, unsigned i, unsigned j, unsigned k)
{
	return (((size_t)n * s->N[0] + i) * s->N[1] + j) * s->N[2] + k;
}
static inline size_t orc_f4idx(const orc_sim* s, int n, unsigned i, unsigned j, unsigned k)
{
	return ((size_t)n + 3 * ((size_t)(k % s->nv) + (size_t)s->nv * (j + (size_t)s->N[1] * i))) * 4 + k / s->nv;
}
static inline float* orc_vref(const orc_sim* s, int n, const unsigned p[3])
{
	return s->sse ? &s->f4_volt[orc_f4idx(s, n, p[0], p[1], p[2])] : &s->volt[orc_idx(s, n, p[0], p[1], p[2])];
}
static inline float* orc_cref(const orc_sim* s, int n, const unsigned p[3])
{
	return s->sse ? &s->f4_curr[orc_f4idx(s, n, p[0], p[1], p[2])] : &s->curr[orc_idx(s, n, p[0], p[1], p[2])];
}
/* tools/useful.cpp:45-75 AssignJobs2Threads: [start, start+num) of thread tid */
static inline void orc_jobs(unsigned jobs, int nth, int tid, unsigned* start, unsigned* num)
{
	unsigned per = jobs / (unsigned)nth, rem = jobs - per * (unsigned)nth, st = 0;
	for (int t = 0; t < tid; ++t) st += per + ((unsigned)t < rem ? 1 : 0);
	*start = st;
	*num = per + ((unsigned)tid < rem ? 1 : 0);
}
#endif
