/*
 * fdtd_oracle_sse.c -- restatement of the reference's sse-compressed + multithreaded engine.
 * TEST INFRASTRUCTURE ONLY (see fdtd_oracle.h).
 *
 * Role: (1) second leg of the cross-engine bit-equality rule (TESTSUITE/enginetests/cavity.m:155)
 * that pins the scalar oracle; (2) the CPU throughput baseline ("reference algorithm,
 * restated") that bench.py times beside the GPU engine.
 *
 * Follows: field layout ArrayENG<f4vector> I-J-K-N with z split into 4 lanes
 * (FDTD/engine_sse.h:38-42, tools/arraylib/array_e.h:57-60); operator de-duplication per f4
 * vector (FDTD/operator_sse_compressed.cpp:114-175); stencil loops
 * (FDTD/engine_sse_compressed.cpp:51-311); x-slab threads with one barrier per phase and per
 * extension hook (FDTD/engine_multithread.cpp:158-197,234-293,310-402); job split
 * (tools/useful.cpp:45-75); FTZ/DAZ (tools/denormal.h:19-30).
 */
#include "fdtd_oracle_priv.h"

#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <emmintrin.h>

typedef float v4sf __attribute__((vector_size(16)));
typedef union { v4sf v; float f[4]; } f4vector; /* tools/array_ops.h:38-44 */

typedef struct {
	orc_sim* s;
	int nth;
	unsigned nv;
	unsigned unique;
	unsigned* op_index;      /* [x][y][z'] */
	f4vector* tab[4][3];     /* vv vi ii iv per component, [unique] */
	pthread_barrier_t bar;
	unsigned iter_ts;
	unsigned *start, *stop;  /* x slab per thread */
} sse_eng;

typedef struct { sse_eng* e; int tid; } targ;

/* ---- Operator_SSE_Compressed::CompressOperator operator_sse_compressed.cpp:114-175 */
typedef struct { f4vector c[12]; } sse_coeff; /* key: 12 f4 vectors, compared with memcmp (:197-200) */

static uint64_t hash_bytes(const void* p, size_t n)
{
	const unsigned char* b = p;
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

static void compress(sse_eng* E)
{
	orc_sim* s = E->s;
	const unsigned Nx = s->N[0], Ny = s->N[1], Nz = s->N[2], nv = E->nv;
	size_t nvec = (size_t)Nx * Ny * nv;
	E->op_index = calloc(nvec, sizeof(unsigned));
	size_t cap = 1024, hcap = 4096;
	sse_coeff* keys = malloc(cap * sizeof(sse_coeff));
	int64_t* htab = malloc(hcap * sizeof(int64_t));
	for (size_t i = 0; i < hcap; ++i) htab[i] = -1;
	unsigned nuniq = 0;
	const float* src[4] = {s->vv, s->vi, s->ii, s->iv};
	for (unsigned i = 0; i < Nx; ++i)
		for (unsigned j = 0; j < Ny; ++j)
			for (unsigned zp = 0; zp < nv; ++zp) {
				sse_coeff k;
				for (int a = 0; a < 4; ++a)
					for (int n = 0; n < 3; ++n)
						for (int l = 0; l < 4; ++l) {
							unsigned z = l * nv + zp;
							/* padding beyond Nz holds zero coefficients (operator_sse.cpp:90-101) */
							k.c[a * 3 + n].f[l] = z < Nz ? src[a][orc_idx(s, n, i, j, z)] : 0.0f;
						}
				uint64_t h = hash_bytes(&k, sizeof(k));
				size_t slot = h & (hcap - 1);
				int64_t found = -1;
				while (htab[slot] >= 0) {
					if (memcmp(&keys[htab[slot]], &k, sizeof(k)) == 0) { found = htab[slot]; break; }
					slot = (slot + 1) & (hcap - 1);
				}
				if (found < 0) {
					if (nuniq == cap) { cap *= 2; keys = realloc(keys, cap * sizeof(sse_coeff)); }
					keys[nuniq] = k;
					htab[slot] = nuniq;
					found = nuniq++;
					if ((size_t)nuniq * 2 > hcap) { /* grow + rehash */
						size_t nh = hcap * 4;
						int64_t* t2 = malloc(nh * sizeof(int64_t));
						for (size_t q = 0; q < nh; ++q) t2[q] = -1;
						for (unsigned u = 0; u < nuniq; ++u) {
							size_t sl = hash_bytes(&keys[u], sizeof(sse_coeff)) & (nh - 1);
							while (t2[sl] >= 0) sl = (sl + 1) & (nh - 1);
							t2[sl] = u;
						}
						free(htab); htab = t2; hcap = nh;
					}
				}
				E->op_index[((size_t)i * Ny + j) * nv + zp] = (unsigned)found;
			}
	E->unique = nuniq;
	for (int a = 0; a < 4; ++a)
		for (int n = 0; n < 3; ++n) {
			if (posix_memalign((void**)&E->tab[a][n], 16, (size_t)(nuniq ? nuniq : 1) * sizeof(f4vector))) abort();
			for (unsigned u = 0; u < nuniq; ++u) E->tab[a][n][u] = keys[u].c[a * 3 + n];
		}
	free(keys); free(htab);
}

/* linear f4 index of component 0 at (x,y,z'): strides N=1, z'=3, y=3*nv, x=3*nv*Ny */
#define F4POS(E, x, y, zp) (3 * ((size_t)(zp) + (size_t)(E)->nv * ((y) + (size_t)(E)->s->N[1] * (x))))

/* Engine_SSE_Compressed::UpdateVoltages engine_sse_compressed.cpp:51-180 */
static void update_voltages(sse_eng* E, unsigned startX, unsigned numX)
{
	orc_sim* s = E->s;
	f4vector* volt = (f4vector*)s->f4_volt;
	f4vector* curr = (f4vector*)s->f4_curr;
	const unsigned Ny = s->N[1], nv = E->nv;
	const long sz = 3, sy = 3 * (long)nv, sx = 3 * (long)nv * Ny;
	f4vector temp;
	for (unsigned x = startX; x < startX + numX; ++x) {
		long shx = (x > 0) * sx;
		for (unsigned y = 0; y < Ny; ++y) {
			long shy = (y > 0) * sy;
			for (unsigned zp = 1; zp < nv; ++zp) {
				unsigned index = E->op_index[((size_t)x * Ny + y) * nv + zp];
				size_t p = F4POS(E, x, y, zp);
				volt[p].v *= E->tab[0][0][index].v;
				volt[p].v += E->tab[1][0][index].v * (curr[p + 2].v - curr[p + 2 - shy].v - curr[p + 1].v + curr[p + 1 - sz].v);
				volt[p + 1].v *= E->tab[0][1][index].v;
				volt[p + 1].v += E->tab[1][1][index].v * (curr[p].v - curr[p - sz].v - curr[p + 2].v + curr[p + 2 - shx].v);
				volt[p + 2].v *= E->tab[0][2][index].v;
				volt[p + 2].v += E->tab[1][2][index].v * (curr[p + 1].v - curr[p + 1 - shx].v - curr[p].v + curr[p - shy].v);
			}
			/* z' = 0: the z-1 neighbour is lane-shifted from z' = nv-1 (:120-176) */
			size_t p0 = F4POS(E, x, y, 0), pe = F4POS(E, x, y, nv - 1);
			unsigned index = E->op_index[((size_t)x * Ny + y) * nv];
			temp.v = (v4sf)_mm_slli_si128((__m128i)curr[pe + 1].v, 4);
			volt[p0].v *= E->tab[0][0][index].v;
			volt[p0].v += E->tab[1][0][index].v * (curr[p0 + 2].v - curr[p0 + 2 - shy].v - curr[p0 + 1].v + temp.v);
			temp.v = (v4sf)_mm_slli_si128((__m128i)curr[pe].v, 4);
			volt[p0 + 1].v *= E->tab[0][1][index].v;
			volt[p0 + 1].v += E->tab[1][1][index].v * (curr[p0].v - temp.v - curr[p0 + 2].v + curr[p0 + 2 - shx].v);
			volt[p0 + 2].v *= E->tab[0][2][index].v;
			volt[p0 + 2].v += E->tab[1][2][index].v * (curr[p0 + 1].v - curr[p0 + 1 - shx].v - curr[p0].v + curr[p0 - shy].v);
		}
	}
}

/* Engine_SSE_Compressed::UpdateCurrents engine_sse_compressed.cpp:182-311 */
static void update_currents(sse_eng* E, unsigned startX, unsigned numX)
{
	orc_sim* s = E->s;
	f4vector* volt = (f4vector*)s->f4_volt;
	f4vector* curr = (f4vector*)s->f4_curr;
	const unsigned Ny = s->N[1], nv = E->nv;
	const long sz = 3, sy = 3 * (long)nv, sx = 3 * (long)nv * Ny;
	f4vector temp;
	for (unsigned x = startX; x < startX + numX; ++x) {
		for (unsigned y = 0; y < Ny - 1; ++y) {
			for (unsigned zp = 0; zp < nv - 1; ++zp) {
				unsigned index = E->op_index[((size_t)x * Ny + y) * nv + zp];
				size_t p = F4POS(E, x, y, zp);
				curr[p].v *= E->tab[2][0][index].v;
				curr[p].v += E->tab[3][0][index].v * (volt[p + 2].v - volt[p + 2 + sy].v - volt[p + 1].v + volt[p + 1 + sz].v);
				curr[p + 1].v *= E->tab[2][1][index].v;
				curr[p + 1].v += E->tab[3][1][index].v * (volt[p].v - volt[p + sz].v - volt[p + 2].v + volt[p + 2 + sx].v);
				curr[p + 2].v *= E->tab[2][2][index].v;
				curr[p + 2].v += E->tab[3][2][index].v * (volt[p + 1].v - volt[p + 1 + sx].v - volt[p].v + volt[p + sy].v);
			}
			/* z' = nv-1: the z+1 neighbour is lane-shifted from z' = 0 (:236-306) */
			size_t p0 = F4POS(E, x, y, 0), pe = F4POS(E, x, y, nv - 1);
			unsigned index = E->op_index[((size_t)x * Ny + y) * nv + nv - 1];
			temp.v = (v4sf)_mm_srli_si128((__m128i)volt[p0 + 1].v, 4);
			curr[pe].v *= E->tab[2][0][index].v;
			curr[pe].v += E->tab[3][0][index].v * (volt[pe + 2].v - volt[pe + 2 + sy].v - volt[pe + 1].v + temp.v);
			temp.v = (v4sf)_mm_srli_si128((__m128i)volt[p0].v, 4);
			curr[pe + 1].v *= E->tab[2][1][index].v;
			curr[pe + 1].v += E->tab[3][1][index].v * (volt[pe].v - temp.v - volt[pe + 2].v + volt[pe + 2 + sx].v);
			curr[pe + 2].v *= E->tab[2][2][index].v;
			curr[pe + 2].v += E->tab[3][2][index].v * (volt[pe + 1].v - volt[pe + 1 + sx].v - volt[pe].v + volt[pe + sy].v);
		}
	}
}

/* NS_Engine_Multithread::thread::operator() engine_multithread.cpp:310-402 */
static void* worker(void* arg)
{
	targ* a = arg;
	sse_eng* E = a->e;
	orc_sim* s = E->s;
	int tid = a->tid, nth = E->nth;
	unsigned old_csr = _mm_getcsr();
	_mm_setcsr(old_csr | 0x8040); /* Denormal::Disable */
	unsigned start = E->start[tid], stop = E->stop[tid];
	unsigned stop_h = (tid == nth - 1) ? stop - 1 : stop; /* engine_multithread.cpp:183-189 */
	for (unsigned it = 0; it < E->iter_ts; ++it) {
		for (int n = s->nexts - 1; n >= 0; --n) { /* Engine_Multithread::DoPreVoltageUpdates :234-243 */
			if (s->exts[n].preV) s->exts[n].preV(s, &s->exts[n], tid, nth);
			pthread_barrier_wait(&E->bar);
		}
		update_voltages(E, start, stop - start + 1);
		pthread_barrier_wait(&E->bar);
		for (int n = 0; n < s->nexts; ++n) {
			if (s->exts[n].postV) s->exts[n].postV(s, &s->exts[n], tid, nth);
			pthread_barrier_wait(&E->bar);
		}
		for (int n = 0; n < s->nexts; ++n) {
			if (s->exts[n].applyV) s->exts[n].applyV(s, &s->exts[n], tid, nth);
			pthread_barrier_wait(&E->bar);
		}
		for (int n = s->nexts - 1; n >= 0; --n) {
			if (s->exts[n].preI) s->exts[n].preI(s, &s->exts[n], tid, nth);
			pthread_barrier_wait(&E->bar);
		}
		if (stop_h + 1 > start) update_currents(E, start, stop_h - start + 1);
		pthread_barrier_wait(&E->bar);
		for (int n = 0; n < s->nexts; ++n) {
			if (s->exts[n].postI) s->exts[n].postI(s, &s->exts[n], tid, nth);
			pthread_barrier_wait(&E->bar);
		}
		for (int n = 0; n < s->nexts; ++n) {
			if (s->exts[n].applyI) s->exts[n].applyI(s, &s->exts[n], tid, nth);
			pthread_barrier_wait(&E->bar);
		}
		if (tid == 0) ++s->numTS;
		pthread_barrier_wait(&E->bar); /* added: makes the numTS read of the next pre-hooks race free */
	}
	_mm_setcsr(old_csr);
	return NULL;
}

void* orc_sse_create(orc_sim* s, int threads)
{
	if (!s || !s->built || s->sse) return NULL;
	if (threads < 1) threads = 1;
	if ((unsigned)threads > s->N[0] - 1) threads = (int)s->N[0] - 1;
	sse_eng* E = calloc(1, sizeof(*E));
	E->s = s; E->nth = threads;
	E->nv = (s->N[2] + 3) / 4; /* ceil(Nz/4), engine_sse.cpp:36 */
	s->nv = E->nv;
	compress(E);
	size_t nf = (size_t)3 * s->N[0] * s->N[1] * E->nv * 4;
	if (posix_memalign((void**)&s->f4_volt, 16, nf * sizeof(float))) abort();
	if (posix_memalign((void**)&s->f4_curr, 16, nf * sizeof(float))) abort();
	memset(s->f4_volt, 0, nf * sizeof(float));
	memset(s->f4_curr, 0, nf * sizeof(float));
	/* carry over the scalar engine's current state */
	s->sse = 1;
	for (int n = 0; n < 3; ++n)
		for (unsigned i = 0; i < s->N[0]; ++i)
			for (unsigned j = 0; j < s->N[1]; ++j)
				for (unsigned k = 0; k < s->N[2]; ++k) {
					s->f4_volt[orc_f4idx(s, n, i, j, k)] = s->volt[orc_idx(s, n, i, j, k)];
					s->f4_curr[orc_f4idx(s, n, i, j, k)] = s->curr[orc_idx(s, n, i, j, k)];
				}
	/* Operator_Multithread::CalcStartStopLines operator_multithread.cpp:93-114 */
	E->start = calloc(threads, sizeof(unsigned));
	E->stop = calloc(threads, sizeof(unsigned));
	for (int t = 0; t < threads; ++t) {
		unsigned st, num;
		orc_jobs(s->N[0], threads, t, &st, &num);
		E->start[t] = st;
		E->stop[t] = st + num - 1;
	}
	return E;
}

void orc_sse_destroy(void* h)
{
	sse_eng* E = h;
	if (!E) return;
	orc_sim* s = E->s;
	free(s->f4_volt); free(s->f4_curr);
	s->f4_volt = s->f4_curr = NULL;
	s->sse = 0;
	for (int a = 0; a < 4; ++a)
		for (int n = 0; n < 3; ++n) free(E->tab[a][n]);
	free(E->op_index); free(E->start); free(E->stop);
	free(E);
}

void orc_sse_iterate(void* h, unsigned n_ts)
{
	sse_eng* E = h;
	E->iter_ts = n_ts;
	pthread_barrier_init(&E->bar, NULL, (unsigned)E->nth);
	pthread_t* th = calloc(E->nth, sizeof(pthread_t));
	targ* args = calloc(E->nth, sizeof(targ));
	for (int t = 0; t < E->nth; ++t) {
		args[t].e = E; args[t].tid = t;
		pthread_create(&th[t], NULL, worker, &args[t]);
	}
	for (int t = 0; t < E->nth; ++t) pthread_join(th[t], NULL);
	pthread_barrier_destroy(&E->bar);
	free(th); free(args);
}

unsigned orc_sse_unique(void* h) { return ((sse_eng*)h)->unique; }
unsigned orc_sse_num_ts(void* h) { return ((sse_eng*)h)->s->numTS; }

void orc_sse_get_fields(void* h, float* volt, float* curr)
{
	sse_eng* E = h;
	orc_sim* s = E->s;
	for (int n = 0; n < 3; ++n)
		for (unsigned i = 0; i < s->N[0]; ++i)
			for (unsigned j = 0; j < s->N[1]; ++j)
				for (unsigned k = 0; k < s->N[2]; ++k) {
					volt[orc_idx(s, n, i, j, k)] = s->f4_volt[orc_f4idx(s, n, i, j, k)];
					curr[orc_idx(s, n, i, j, k)] = s->f4_curr[orc_f4idx(s, n, i, j, k)];
				}
}
