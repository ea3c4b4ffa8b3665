/* TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
 *
 * Drives the reference's OWN classes -- compiled unmodified from /root/reference by oracle/Makefile.ref --
 * through the setup sequence of openEMS::SetupFDTD (openems.cpp:1127-1344) and exposes the result through a
 * C API that mirrors oracle/fdtd_oracle.h one to one (ref_* instead of orc_*), so that the same test cases
 * can be run through the CPU restatement and through the real reference code and compared bit for bit.
 * openems.cpp itself (XML, command line, HDF5/VTK) is the control plane and is not compiled.
 * This file is built with -fno-access-control: it reads protected members of the reference classes
 * (extension coefficient tables, engine-extension state) instead of editing their headers.
 */
#include "FDTD/operator.h"
#include "FDTD/operator_sse.h"
#include "FDTD/operator_sse_compressed.h"
#include "FDTD/operator_multithread.h"
#include "FDTD/engine.h"
#include "FDTD/engine_sse.h"
#include "FDTD/engine_sse_compressed.h"
#include "FDTD/engine_multithread.h"
#include "FDTD/excitation.h"
#include "FDTD/engine_interface_fdtd.h"
#include "FDTD/engine_interface_sse_fdtd.h"
#include "FDTD/extensions/operator_ext_excitation.h"
#include "FDTD/extensions/operator_ext_tfsf.h"
#include "FDTD/extensions/operator_ext_mur_abc.h"
#include "FDTD/extensions/operator_ext_upml.h"
#include "FDTD/extensions/operator_ext_lorentzmaterial.h"
#include "FDTD/extensions/operator_ext_conductingsheet.h"
#include "FDTD/extensions/operator_ext_lumpedRLC.h"
#include "FDTD/extensions/operator_ext_steadystate.h"
#include "FDTD/extensions/operator_ext_absorbing_bc.h"
#include "FDTD/extensions/engine_ext_upml.h"
#include "FDTD/extensions/engine_ext_mur_abc.h"
#include "FDTD/extensions/engine_ext_steadystate.h"
#include "FDTD/extensions/engine_ext_absorbing_bc.h"
#include "Common/processvoltage.h"
#include "Common/processcurrent.h"
#include "Common/processfieldprobe.h"
#include "Common/processfields_td.h"
#include "Common/processfields_fd.h"
#include "Common/processmodematch.h"
#include "tools/denormal.h"
#include "ref_recorder.h"
#include "ref_driver.h"

#include "ContinuousStructure.h"
#include "CSPrimBox.h"

#include <cstring>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <complex>

#ifdef REF_WITH_CUDA
#include "operator_cuda.h"
#include "engine_cuda.h"
#include "engine_interface_cuda_fdtd.h"
#include "processing_cuda.h"
#endif

struct ref_rlc_raw {
	unsigned count;
	std::vector<int> dir;
	std::vector<unsigned> pos;
	std::vector<float> c[9];
	Operator_Ext_LumpedRLC* ext;
};

struct ref_sim {
	ContinuousStructure* csx;
	Excitation* exc;
	Operator* op;
	Engine* eng;
	Engine_Interface_FDTD* eif;
	ProcessingArray* PA;
	int engine_kind, threads;
	int fast_processing;   // with the CUDA engine: use the Process*_CUDA classes (integration/processing_cuda.h) instead of the stock ones
	unsigned N[3];
	std::vector<double> lines[3];
	int bc[6];
	unsigned pml[6];
	double mur_v[6];
	double forced_dT, ts_factor;
	int cell_constant_material;
	bool built;
	// steady state request
	unsigned ss_period_ts;
	std::vector<unsigned> ss_pos[3];
	std::vector<int> ss_dir;
	Operator_Ext_SteadyState* ss_op;
	// TFSF request (mesh indices)
	bool tfsf_on;
	std::vector<ref_rlc_raw> rlc_raw;
	// extension pointers after build
	std::vector<Operator_Ext_UPML*> upml;
	std::vector<Operator_Ext_Mur_ABC*> mur;
	std::vector<Operator_Ext_Absorbing_BC*> abc;
	Operator_Ext_LorentzMaterial* lor;
	std::vector<Operator_Ext_LorentzMaterial*> lor_all; // Lorentz first, then conducting sheet
	Operator_Ext_Excitation* exc_ext;
	Operator_Ext_TFSF* tfsf;
	// scratch buffers handed out to the caller
	std::vector<float> buf_coeff[4], buf_volt, buf_curr;
	std::vector<std::vector<float> > buf_upml, buf_flux;
	std::vector<std::vector<float> > buf_mur;
	std::vector<std::vector<float> > buf_misc;
	std::vector<Processing*> procs;
};

// the reference runs with FTZ/DAZ set (tools/denormal.h:19-30, called from openEMS::SetupFDTD / engine threads);
// set it for the duration of a call only, so that the host process (numpy) keeps its own MXCSR
struct ftz_scope {
	unsigned int saved;
	ftz_scope() { saved = _mm_getcsr(); Denormal::Disable(); }
	~ftz_scope() { _mm_setcsr(saved); }
};

static size_t ncell(const ref_sim* s) { return (size_t)s->N[0] * s->N[1] * s->N[2]; }

// the engine extension created from a given operator extension (most CreateEngineExtention() implementations
// do not store it in m_Eng_Ext, so search the engine's list: engine.cpp:75-80)
template <class EngExt>
static EngExt* find_eng_ext(const ref_sim* s, const Operator_Extension* op_ext)
{
	if (!s->eng) return NULL;
	for (size_t n = 0; n < s->eng->GetExtensionCount(); ++n) {
		EngExt* e = dynamic_cast<EngExt*>(s->eng->GetExtension(n));
		if (e && e->m_Op_ext == op_ext) return e;
	}
	return NULL;
}

static void segv_backtrace(int sig)
{
	void* frames[64];
	int n = backtrace(frames, 64);
	backtrace_symbols_fd(frames, n, 2);
	signal(sig, SIG_DFL);
	raise(sig);
}

extern "C" {

ref_sim* ref_create(const unsigned nlines[3], const double* x, const double* y, const double* z, double grid_delta)
{
	if (nlines[0] < 3 || nlines[1] < 3 || nlines[2] < 3) return NULL;
	if (getenv("REF_DEBUG")) signal(SIGSEGV, segv_backtrace);
	ref_sim* s = new ref_sim();
	s->csx = new ContinuousStructure();
	const double* l[3] = { x, y, z };
	for (int n = 0; n < 3; ++n) {
		s->N[n] = nlines[n];
		s->lines[n].assign(l[n], l[n] + nlines[n]);
		s->csx->GetGrid()->AddDiscLines(n, nlines[n], l[n]);
	}
	s->csx->GetGrid()->SetDeltaUnit(grid_delta);
	s->exc = new Excitation();
	s->op = NULL; s->eng = NULL; s->eif = NULL; s->PA = NULL;
	s->engine_kind = 0; s->threads = 1; s->fast_processing = 1;
	for (int n = 0; n < 6; ++n) { s->bc[n] = 0; s->pml[n] = 8; s->mur_v[n] = -1; }
	s->forced_dT = 0; s->ts_factor = 1; s->cell_constant_material = 0;
	s->built = false;
	s->ss_period_ts = 0; s->ss_op = NULL;
	s->tfsf_on = false;
	s->lor = NULL; s->exc_ext = NULL; s->tfsf = NULL;
	return s;
}

void ref_destroy(ref_sim* s)
{
	if (!s) return;
	const bool dbg = getenv("REF_DEBUG") != NULL;
#define DBG(msg) do { if (dbg) { fprintf(stderr, "ref_destroy: %s\n", msg); fflush(stderr); } } while (0)
	DBG("processings");
	if (s->PA) s->PA->DeleteAll();
	delete s->PA;
	DBG("engine interface");
	delete s->eif;
	// openEMS deletes the engine before the operator (openems.cpp:99-113)
	DBG("engine");
	delete s->eng;
	DBG("operator");
	delete s->op;
	DBG("excitation");
	delete s->exc;
	DBG("csx");
	delete s->csx;
	DBG("done");
	delete s;
#undef DBG
}

/* 0 basic (Engine), 1 sse (Engine_sse), 2 sse-compressed, 3 multithreaded (openems.cpp:224-251,738-753), 4 cuda */
void ref_set_engine(ref_sim* s, int kind, int threads) { s->engine_kind = kind; s->threads = threads; }
void ref_set_fast_processing(ref_sim* s, int on) { s->fast_processing = on; }
int ref_has_cuda(void)
{
#ifdef REF_WITH_CUDA
	return 1;
#else
	return 0;
#endif
}

void ref_set_bc(ref_sim* s, const int bc[6], const unsigned pml_size[6])
{
	for (int n = 0; n < 6; ++n) { s->bc[n] = bc[n]; s->pml[n] = pml_size[n]; }
}
void ref_set_background(ref_sim* s, double epsR, double mueR, double kappa, double sigma)
{
	CSBackgroundMaterial* bg = s->csx->GetBackgroundMaterial();
	bg->SetEpsilon(epsR); bg->SetMue(mueR); bg->SetKappa(kappa); bg->SetSigma(sigma);
}
void ref_set_mur_phase_velocity(ref_sim* s, double v) { for (int n = 0; n < 6; ++n) s->mur_v[n] = v; }
void ref_set_timestep(ref_sim* s, double forced_dT, double factor) { s->forced_dT = forced_dT; s->ts_factor = factor; }

static int add_box(ref_sim* s, CSProperties* prop, int prio, const double start[3], const double stop[3])
{
	s->csx->AddProperty(prop);
	CSPrimBox* box = new CSPrimBox(prop, start, stop, prio);
	s->csx->RegisterPrimitive(box);
	return (int)prop->GetID();
}

int ref_add_material(ref_sim* s, int prio, const double start[3], const double stop[3], double epsR, double mueR, double kappa, double sigma)
{
	CSPropMaterial* m = new CSPropMaterial();
	m->SetEpsilon(epsR); m->SetMue(mueR); m->SetKappa(kappa); m->SetSigma(sigma);
	return add_box(s, m, prio, start, stop);
}
int ref_add_metal(ref_sim* s, int prio, const double start[3], const double stop[3])
{
	return add_box(s, new CSPropMetal(), prio, start, stop);
}
int ref_add_lorentz(ref_sim* s, int prio, const double start[3], const double stop[3], double epsR, double mueR, double kappa, double sigma,
                    int order, const double* eps_fp, const double* eps_tau, const double* eps_flor,
                    const double* mue_fp, const double* mue_tau, const double* mue_flor)
{
	CSPropLorentzMaterial* m = new CSPropLorentzMaterial();
	m->SetEpsilon(epsR); m->SetMue(mueR); m->SetKappa(kappa); m->SetSigma(sigma);
	m->SetDispersionOrder(order);
	for (int o = 0; o < order; ++o) {
		m->SetEpsPlasmaFreq(o, eps_fp[o]); m->SetEpsRelaxTime(o, eps_tau[o]); m->SetEpsLorPoleFreq(o, eps_flor[o]);
		m->SetMuePlasmaFreq(o, mue_fp[o]); m->SetMueRelaxTime(o, mue_tau[o]); m->SetMueLorPoleFreq(o, mue_flor[o]);
	}
	return add_box(s, m, prio, start, stop);
}
int ref_add_debye(ref_sim* s, int prio, const double start[3], const double stop[3], double epsR, double mueR, double kappa, double sigma,
                  int order, const double* eps_delta, const double* eps_tau)
{
	CSPropDebyeMaterial* m = new CSPropDebyeMaterial();
	m->SetEpsilon(epsR); m->SetMue(mueR); m->SetKappa(kappa); m->SetSigma(sigma);
	m->SetDispersionOrder(order);
	for (int o = 0; o < order; ++o) { m->SetEpsDelta(o, eps_delta[o]); m->SetEpsRelaxTime(o, eps_tau[o]); }
	return add_box(s, m, prio, start, stop);
}
int ref_add_conducting_sheet(ref_sim* s, int prio, const double start[3], const double stop[3], double conductivity, double thickness)
{
	return add_box(s, new CSPropConductingSheet(conductivity, thickness), prio, start, stop);
}
int ref_add_excitation(ref_sim* s, int prio, const double start[3], const double stop[3], int exc_type, const double vec[3], double delay_s)
{
	CSPropExcitation* e = new CSPropExcitation();
	e->SetExcitType(exc_type);
	for (int n = 0; n < 3; ++n) e->SetExcitation(vec[n], n);
	e->SetDelay(delay_s);
	return add_box(s, e, prio, start, stop);
}
int ref_add_lumped_rc(ref_sim* s, const double start[3], const double stop[3], int dir, double R, double C, int caps)
{
	CSPropLumpedElement* e = new CSPropLumpedElement();
	e->SetResistance(R); e->SetCapacity(C); e->SetDirection(dir); e->SetCaps(caps != 0);
	e->SetLEtype(CSPropLumpedElement::PARALLEL);
	return add_box(s, e, 0, start, stop);
}
/* series (type 1) or parallel (type 0) RLC through the reference's own coefficient builder
   (FDTD/extensions/operator_ext_lumpedRLC.cpp:112-534) */
int ref_add_lumped_rlc(ref_sim* s, const double start[3], const double stop[3], int dir, double R, double C, double L, int type, int caps)
{
	CSPropLumpedElement* e = new CSPropLumpedElement();
	e->SetResistance(R); e->SetCapacity(C); e->SetInductance(L); e->SetDirection(dir); e->SetCaps(caps != 0);
	e->SetLEtype(type == 1 ? CSPropLumpedElement::SERIES : CSPropLumpedElement::PARALLEL);
	return add_box(s, e, 0, start, stop);
}
int ref_add_rlc_raw(ref_sim* s, unsigned count, const int* dir, const unsigned* pos, const float* ilv, const float* i2v, const float* vvd,
                    const float* vv2, const float* vj1, const float* vj2, const float* ib0, const float* b1, const float* b2)
{
	ref_rlc_raw r;
	r.count = count;
	r.dir.assign(dir, dir + count);
	r.pos.assign(pos, pos + 3 * count);
	const float* c[9] = { ilv, i2v, vvd, vv2, vj1, vj2, ib0, b1, b2 };
	for (int i = 0; i < 9; ++i) r.c[i].assign(c[i], c[i] + count);
	r.ext = NULL;
	s->rlc_raw.push_back(r);
	return (int)s->rlc_raw.size() - 1;
}
int ref_add_steadystate(ref_sim* s, unsigned period_ts, unsigned count, const unsigned* pos3, const int* dir)
{
	s->ss_period_ts = period_ts;
	for (int n = 0; n < 3; ++n) s->ss_pos[n].assign(pos3 + n * count, pos3 + (n + 1) * count);
	s->ss_dir.assign(dir, dir + count);
	return 0;
}
double ref_steadystate_last_diff(const ref_sim* s)
{
	if (!s->ss_op) return -1;
	Engine_Ext_SteadyState* e = dynamic_cast<Engine_Ext_SteadyState*>(s->ss_op->GetEngineExtention());
	return e ? e->GetLastDiff() : -1;
}
int ref_set_tfsf(ref_sim* s, const unsigned start[3], const unsigned stop[3], const double prop_dir[3], const double e_amp[3])
{
	CSPropExcitation* e = new CSPropExcitation();
	e->SetExcitType(10);
	for (int n = 0; n < 3; ++n) { e->SetExcitation(e_amp[n], n); e->SetPropagationDir(prop_dir[n], n); }
	double c0[3], c1[3];
	for (int n = 0; n < 3; ++n) { c0[n] = s->lines[n][start[n]]; c1[n] = s->lines[n][stop[n]]; }
	add_box(s, e, 0, c0, c1);
	s->tfsf_on = true;
	return 0;
}
int ref_add_absorbing_sheet(ref_sim* s, const unsigned x0[3], const unsigned x1[3], int normal_positive, int type, double phase_velocity)
{
	CSPropAbsorbingBC* p = new CSPropAbsorbingBC();
	p->SetNormalSignPositive(normal_positive != 0);
	p->SetAbsorbingBoundaryType(type);
	p->SetPhaseVelocity(phase_velocity);
	double c0[3], c1[3];
	for (int n = 0; n < 3; ++n) { c0[n] = s->lines[n][x0[n]]; c1[n] = s->lines[n][x1[n]]; }
	add_box(s, p, 0, c0, c1);
	return 0;
}

void ref_set_excite_gauss(ref_sim* s, double f0, double fc) { s->exc->SetupGaussianPulse(f0, fc); }
void ref_set_excite_sinus(ref_sim* s, double f0) { s->exc->SetupSinusoidal(f0); }
void ref_set_excite_dirac(ref_sim* s, double fmax) { s->exc->SetupDiracPulse(fmax); }
void ref_set_excite_step(ref_sim* s, double fmax) { s->exc->SetupStepExcite(fmax); }
void ref_set_excite_custom(ref_sim* s, const char* func, double f0, double fmax) { s->exc->SetupCustomExcite(func, f0, fmax); }

/* openEMS::SetupFDTD openems.cpp:1161-1316 (operator part), SetupBoundaryConditions :383-408, SetupAbsorbingSheets :410-445 */
int ref_build(ref_sim* s, unsigned max_ts)
{
	if (s->built) return -1;
	ftz_scope ftz;
	switch (s->engine_kind) {   // openEMS::SetupOperator openems.cpp:738-753
	case 1: s->op = Operator_sse::New(); break;
	case 2: s->op = Operator_SSE_Compressed::New(); break;
	case 3: s->op = Operator_Multithread::New(s->threads); break;
#ifdef REF_WITH_CUDA
	case 4: s->op = Operator_CUDA::New(); break;
#endif
	default: s->op = Operator::New(); break;
	}
	Operator* op = s->op;
	op->SetQuarterCellMaterialAvg();
	if (s->cell_constant_material) op->SetCellConstantMaterial();
	op->SetExcitationSignal(s->exc);
	op->AddExtension(new Operator_Ext_Excitation(op));
	op->AddExtension(new Operator_Ext_TFSF(op));
	if (!op->SetGeometryCSX(s->csx)) return -2;

	// SetupBoundaryConditions
	op->SetBoundaryCondition(s->bc);
	for (int n = 0; n < 6; ++n) {
		op->SetBCSize(n, 0);
		if (s->bc[n] == 2) {
			op->SetBCSize(n, 1);
			Operator_Ext_Mur_ABC* m = new Operator_Ext_Mur_ABC(op);
			m->SetDirection(n / 2, n % 2);
			if (s->mur_v[n] > 0) m->SetPhaseVelocity(s->mur_v[n]);
			op->AddExtension(m);
		}
		if (s->bc[n] == 3) op->SetBCSize(n, s->pml[n]);
	}
	Operator_Ext_UPML::Create_UPML(op, s->bc, s->pml, std::string());

	if (s->forced_dT > 0) op->SetTimestep(s->forced_dT);
	if (s->ts_factor < 1) op->SetTimestepFactor(s->ts_factor);

	if (s->ss_period_ts > 0) {
		// the period in seconds is fixed up after the build (the timestep is not known yet)
		s->ss_op = new Operator_Ext_SteadyState(op, 1.0);
		for (size_t i = 0; i < s->ss_dir.size(); ++i) {
			unsigned pos[3] = { s->ss_pos[0][i], s->ss_pos[1][i], s->ss_pos[2][i] };
			s->ss_op->Add_E_Probe(pos, s->ss_dir[i]);
		}
		op->AddExtension(s->ss_op);
	}
	if ((s->csx->GetQtyPropertyType(CSProperties::LORENTZMATERIAL) > 0) || (s->csx->GetQtyPropertyType(CSProperties::DEBYEMATERIAL) > 0))
		op->AddExtension(new Operator_Ext_LorentzMaterial(op));
	if (s->csx->GetQtyPropertyType(CSProperties::CONDUCTINGSHEET) > 0)
		op->AddExtension(new Operator_Ext_ConductingSheet(op, s->exc->GetMaxFreq()));
	if (s->csx->GetQtyPropertyType(CSProperties::LUMPED_ELEMENT) > 0)
		op->AddExtension(new Operator_Ext_LumpedRLC(op));
	for (size_t r = 0; r < s->rlc_raw.size(); ++r) {
		s->rlc_raw[r].ext = new Operator_Ext_LumpedRLC(op);
		op->AddExtension(s->rlc_raw[r].ext);
	}
	if (s->csx->GetQtyPropertyType(CSProperties::ABSORBING_BC) > 0) {
		std::vector<CSProperties*> props = s->csx->GetPropertyByType(CSProperties::ABSORBING_BC);
		for (size_t n = 0; n < props.size(); ++n) {
			CSPropAbsorbingBC* p = dynamic_cast<CSPropAbsorbingBC*>(props[n]);
			std::vector<CSPrimitives*> prims = p->GetAllPrimitives();
			for (size_t i = 0; i < prims.size(); ++i) {
				Operator_Ext_Absorbing_BC* a = new Operator_Ext_Absorbing_BC(op);
				if (a->SetInitParams(prims[i], p)) op->AddExtension(a);
				else delete a;
			}
		}
	}

	if (op->CalcECOperator(Operator::None) != 0) return -3;

	// raw RLC coefficient tables (tests with hand-made coefficients): fill the extension's arrays as
	// Operator_Ext_LumpedRLC::BuildExtension would (operator_ext_lumpedRLC.cpp:494-531)
	for (size_t r = 0; r < s->rlc_raw.size(); ++r) {
		ref_rlc_raw& rr = s->rlc_raw[r];
		Operator_Ext_LumpedRLC* e = rr.ext;
		unsigned cnt = rr.count;
		e->RLC_count = cnt;
		if (!cnt) continue;
		e->v_RLC_dir = new int[cnt];
		FDTD_FLOAT** dst[9] = { &e->v_RLC_ilv, &e->v_RLC_i2v, &e->v_RLC_vvd, &e->v_RLC_vv2, &e->v_RLC_vj1, &e->v_RLC_vj2, &e->v_RLC_ib0, &e->v_RLC_b1, &e->v_RLC_b2 };
		for (int i = 0; i < 9; ++i) { *dst[i] = new FDTD_FLOAT[cnt]; std::copy(rr.c[i].begin(), rr.c[i].end(), *dst[i]); }
		e->v_RLC_pos = new unsigned int*[3];
		for (int n = 0; n < 3; ++n) {
			e->v_RLC_pos[n] = new unsigned int[cnt];
			std::copy(rr.pos.begin() + n * cnt, rr.pos.begin() + (n + 1) * cnt, e->v_RLC_pos[n]);
		}
		std::copy(rr.dir.begin(), rr.dir.end(), e->v_RLC_dir);
	}
	if (s->ss_op) {
		s->ss_op->m_TS_period = s->ss_period_ts;
		s->ss_op->m_T_period = s->ss_period_ts * op->GetTimestep();
	}

	if (!s->exc->buildExcitationSignal(max_ts)) return -4;

	// collect the extensions that survived "remove inactive extensions" (operator.cpp:1063-1074)
	for (size_t n = 0; n < op->GetNumberOfExtentions(); ++n) {
		Operator_Extension* e = op->GetExtension(n);
		if (Operator_Ext_UPML* u = dynamic_cast<Operator_Ext_UPML*>(e)) s->upml.push_back(u);
		else if (Operator_Ext_Mur_ABC* m = dynamic_cast<Operator_Ext_Mur_ABC*>(e)) s->mur.push_back(m);
		else if (Operator_Ext_Absorbing_BC* a = dynamic_cast<Operator_Ext_Absorbing_BC*>(e)) s->abc.push_back(a);
		else if (Operator_Ext_LorentzMaterial* l = dynamic_cast<Operator_Ext_LorentzMaterial*>(e)) { if (!s->lor) s->lor = l; s->lor_all.push_back(l); }
		else if (Operator_Ext_Excitation* x = dynamic_cast<Operator_Ext_Excitation*>(e)) s->exc_ext = x;
		else if (Operator_Ext_TFSF* t = dynamic_cast<Operator_Ext_TFSF*>(e)) s->tfsf = t;
	}

	s->eng = op->CreateEngine();
	if (!s->eng) return -5;
	// openEMS::NewEngineInterface openems.cpp:447-476
#ifdef REF_WITH_CUDA
	if (dynamic_cast<Operator_CUDA*>(op)) s->eif = new Engine_Interface_CUDA_FDTD(dynamic_cast<Operator_CUDA*>(op));
	else
#endif
	if (Operator_sse* os = dynamic_cast<Operator_sse*>(op)) s->eif = new Engine_Interface_SSE_FDTD(os);
	else s->eif = new Engine_Interface_FDTD(op);
	if (s->ss_op) {
		Engine_Ext_SteadyState* e = dynamic_cast<Engine_Ext_SteadyState*>(s->ss_op->GetEngineExtention());
		if (e) {
#ifdef REF_WITH_CUDA
			if (dynamic_cast<Operator_CUDA*>(op)) e->SetEngineInterface(new Engine_Interface_CUDA_FDTD(dynamic_cast<Operator_CUDA*>(op)));
			else
#endif
			if (Operator_sse* os = dynamic_cast<Operator_sse*>(op)) e->SetEngineInterface(new Engine_Interface_SSE_FDTD(os));
			else e->SetEngineInterface(new Engine_Interface_FDTD(op));
		}
	}
	s->PA = new ProcessingArray(s->exc->GetNyquistNum());
	s->built = true;
	return 0;
}

double ref_dT(const ref_sim* s) { return s->op->GetTimestep(); }
unsigned ref_nyquist(const ref_sim* s) { return s->exc->GetNyquistNum(); }
const float* ref_coeff(const ref_sim* cs, int which)
{
	ref_sim* s = const_cast<ref_sim*>(cs);
	std::vector<float>& b = s->buf_coeff[which];
	b.resize(3 * ncell(s));
	size_t p = 0;
	for (unsigned n = 0; n < 3; ++n)
		for (unsigned i = 0; i < s->N[0]; ++i)
			for (unsigned j = 0; j < s->N[1]; ++j)
				for (unsigned k = 0; k < s->N[2]; ++k, ++p)
					b[p] = which == 0 ? s->op->GetVV(n, i, j, k) : which == 1 ? s->op->GetVI(n, i, j, k)
					     : which == 2 ? s->op->GetII(n, i, j, k) : s->op->GetIV(n, i, j, k);
	return b.data();
}
unsigned ref_signal_length(const ref_sim* s) { return s->exc->GetLength(); }
const float* ref_signal(const ref_sim* s, int is_curr) { return is_curr ? s->exc->GetCurrentSignal() : s->exc->GetVoltageSignal(); }
unsigned ref_signal_period_ts(const ref_sim* s)
{
	// Engine_Ext_Excitation engine_ext_excitation.cpp:41-43: p = int(period/dT)
	double per = s->exc->GetSignalPeriod();
	return per > 0 ? (unsigned)(int)(per / s->exc->GetTimestep()) : 0;
}
unsigned ref_exc_count(const ref_sim* s, int is_curr) { return !s->exc_ext ? 0 : (is_curr ? s->exc_ext->Curr_Count : s->exc_ext->Volt_Count); }
void ref_exc_get(const ref_sim* s, int is_curr, unsigned* idx, unsigned* dir, float* amp, unsigned* delay)
{
	Operator_Ext_Excitation* e = s->exc_ext;
	unsigned cnt = ref_exc_count(s, is_curr);
	for (unsigned i = 0; i < cnt; ++i) {
		for (int n = 0; n < 3; ++n) idx[n * cnt + i] = is_curr ? e->Curr_index[n][i] : e->Volt_index[n][i];
		dir[i] = is_curr ? e->Curr_dir[i] : e->Volt_dir[i];
		amp[i] = is_curr ? e->Curr_amp[i] : e->Volt_amp[i];
		delay[i] = is_curr ? e->Curr_delay[i] : e->Volt_delay[i];
	}
}
int ref_upml_count(const ref_sim* s) { return (int)s->upml.size(); }
void ref_upml_box(const ref_sim* s, int b, unsigned start[3], unsigned nlines[3])
{
	for (int n = 0; n < 3; ++n) { start[n] = s->upml[b]->m_StartPos[n]; nlines[n] = s->upml[b]->m_numLines[n]; }
}
static const float* copy_nijk(ref_sim* s, std::vector<std::vector<float> >& store, size_t slot, ArrayLib::ArrayNIJK<FDTD_FLOAT>& a, const unsigned nl[3])
{
	if (store.size() <= slot) store.resize(slot + 1);
	std::vector<float>& b = store[slot];
	b.resize((size_t)3 * nl[0] * nl[1] * nl[2]);
	size_t p = 0;
	for (unsigned n = 0; n < 3; ++n)
		for (unsigned i = 0; i < nl[0]; ++i)
			for (unsigned j = 0; j < nl[1]; ++j)
				for (unsigned k = 0; k < nl[2]; ++k) b[p++] = a(n, i, j, k);
	(void)s;
	return b.data();
}
const float* ref_upml_coeff(const ref_sim* cs, int b, int which)
{
	ref_sim* s = const_cast<ref_sim*>(cs);
	Operator_Ext_UPML* u = s->upml[b];
	ArrayLib::ArrayNIJK<FDTD_FLOAT>* arr[6] = { &u->vv, &u->vvfn, &u->vvfo, &u->ii, &u->iifn, &u->iifo };
	return copy_nijk(s, s->buf_upml, (size_t)b * 6 + which, *arr[which], u->m_numLines);
}
const float* ref_upml_flux(const ref_sim* cs, int b, int is_curr)
{
	ref_sim* s = const_cast<ref_sim*>(cs);
	Engine_Ext_UPML* e = find_eng_ext<Engine_Ext_UPML>(s, s->upml[b]);
	if (!e) return NULL;
	return copy_nijk(s, s->buf_flux, (size_t)b * 2 + (is_curr ? 1 : 0), is_curr ? e->curr_flux : e->volt_flux, s->upml[b]->m_numLines);
}
int ref_mur_count(const ref_sim* s) { return (int)s->mur.size(); }
void ref_mur_info(const ref_sim* s, int m, int* ny, int* top, unsigned* line, unsigned* shift, unsigned nlines[2], unsigned* start_ts)
{
	Operator_Ext_Mur_ABC* o = s->mur[m];
	*ny = o->m_ny; *top = o->m_top; *line = o->m_LineNr; *shift = (unsigned)o->m_LineNr_Shift;
	nlines[0] = o->m_numLines[0]; nlines[1] = o->m_numLines[1];
	Engine_Ext_Mur_ABC* e = find_eng_ext<Engine_Ext_Mur_ABC>(s, o);
	*start_ts = e ? e->m_start_TS : 0;
}
const float* ref_mur_coeff(const ref_sim* cs, int m, int which)
{
	ref_sim* s = const_cast<ref_sim*>(cs);
	Operator_Ext_Mur_ABC* o = s->mur[m];
	size_t slot = (size_t)m * 2 + which;
	if (s->buf_mur.size() <= slot) s->buf_mur.resize(slot + 1);
	std::vector<float>& b = s->buf_mur[slot];
	b.resize((size_t)o->m_numLines[0] * o->m_numLines[1]);
	size_t p = 0;
	for (unsigned i = 0; i < o->m_numLines[0]; ++i)
		for (unsigned j = 0; j < o->m_numLines[1]; ++j)
			b[p++] = which ? o->m_Mur_Coeff_nyPP(i, j) : o->m_Mur_Coeff_nyP(i, j);
	return b.data();
}
/* ext selects the dispersive extension: 0 = Operator_Ext_LorentzMaterial, 1 = next (conducting sheet) ... */
int ref_lorentz_ext_count(const ref_sim* s) { return (int)s->lor_all.size(); }
void ref_lorentz_select(ref_sim* s, int ext) { s->lor = (ext >= 0 && ext < (int)s->lor_all.size()) ? s->lor_all[ext] : NULL; }
int ref_lorentz_order(const ref_sim* s) { return s->lor ? s->lor->m_Order : 0; }
unsigned ref_lorentz_count(const ref_sim* s, int o) { return s->lor ? s->lor->m_LM_Count.at(o) : 0; }
int ref_lorentz_flags(const ref_sim* s, int o)
{
	if (!s->lor) return 0;
	return (s->lor->m_volt_ADE_On[o] ? 1 : 0) | (s->lor->m_curr_ADE_On[o] ? 2 : 0)
	     | (s->lor->m_volt_Lor_ADE_On[o] ? 4 : 0) | (s->lor->m_curr_Lor_ADE_On[o] ? 8 : 0);
}
const unsigned* ref_lorentz_pos(const ref_sim* s, int o, int n) { return s->lor ? s->lor->m_LM_pos[o][n] : NULL; }
const float* ref_lorentz_coeff(const ref_sim* s, int o, int which, int n)
{
	if (!s->lor) return NULL;
	Operator_Ext_LorentzMaterial* l = s->lor;
	bool vOn = l->m_volt_ADE_On[o], cOn = l->m_curr_ADE_On[o], vL = l->m_volt_Lor_ADE_On[o], cL = l->m_curr_Lor_ADE_On[o];
	switch (which) {
	case 0: return vOn ? l->v_int_ADE[o][n] : NULL;
	case 1: return vOn ? l->v_ext_ADE[o][n] : NULL;
	case 2: return vL ? l->v_Lor_ADE[o][n] : NULL;
	case 3: return cOn ? l->i_int_ADE[o][n] : NULL;
	case 4: return cOn ? l->i_ext_ADE[o][n] : NULL;
	case 5: return cL ? l->i_Lor_ADE[o][n] : NULL;
	}
	return NULL;
}
/* lumped RLC tables built by the reference (ext 0 = the CSX-driven extension when present) */
unsigned ref_rlc_count(const ref_sim* s)
{
	for (size_t n = 0; n < s->op->GetNumberOfExtentions(); ++n)
		if (Operator_Ext_LumpedRLC* r = dynamic_cast<Operator_Ext_LumpedRLC*>(s->op->GetExtension(n))) return r->RLC_count;
	return 0;
}
void ref_rlc_get(const ref_sim* s, int* dir, unsigned* pos, float* coeffs9)
{
	for (size_t n = 0; n < s->op->GetNumberOfExtentions(); ++n)
		if (Operator_Ext_LumpedRLC* r = dynamic_cast<Operator_Ext_LumpedRLC*>(s->op->GetExtension(n))) {
			unsigned cnt = r->RLC_count;
			const FDTD_FLOAT* c[9] = { r->v_RLC_ilv, r->v_RLC_i2v, r->v_RLC_vvd, r->v_RLC_vv2, r->v_RLC_vj1, r->v_RLC_vj2, r->v_RLC_ib0, r->v_RLC_b1, r->v_RLC_b2 };
			for (unsigned i = 0; i < cnt; ++i) {
				dir[i] = r->v_RLC_dir[i];
				for (int d = 0; d < 3; ++d) pos[d * cnt + i] = r->v_RLC_pos[d][i];
				for (int k = 0; k < 9; ++k) coeffs9[k * cnt + i] = c[k][i];
			}
			return;
		}
}

void ref_iterate(ref_sim* s, unsigned n_ts) { ftz_scope ftz; if (n_ts) s->eng->IterateTS(n_ts); }
unsigned ref_num_ts(const ref_sim* s) { return s->eng->GetNumberOfTimesteps(); }
static float* snapshot(ref_sim* s, bool curr)
{
	std::vector<float>& b = curr ? s->buf_curr : s->buf_volt;
	b.resize(3 * ncell(s));
#ifdef REF_WITH_CUDA
	if (Engine_CUDA* ec = dynamic_cast<Engine_CUDA*>(s->eng)) {   // one bulk copy instead of 3N per-cell GetVolt round trips
		if (oems_cuda_get_fields(ec->GetHandle(), curr ? 1 : 0, b.data()) == 0) return b.data();
	}
#endif
	size_t p = 0;
	for (unsigned n = 0; n < 3; ++n)
		for (unsigned i = 0; i < s->N[0]; ++i)
			for (unsigned j = 0; j < s->N[1]; ++j)
				for (unsigned k = 0; k < s->N[2]; ++k, ++p)
					b[p] = curr ? s->eng->GetCurr(n, i, j, k) : s->eng->GetVolt(n, i, j, k);
	return b.data();
}
/* snapshots (copies) of the engine's fields in ArrayNIJK order */
float* ref_volt(ref_sim* s) { return snapshot(s, false); }
float* ref_curr(ref_sim* s) { return snapshot(s, true); }
void ref_set_field(ref_sim* s, int is_curr, int n, unsigned x, unsigned y, unsigned z, float v)
{
	if (is_curr) s->eng->SetCurr(n, x, y, z, v); else s->eng->SetVolt(n, x, y, z, v);
}
void ref_reset_fields(ref_sim* s)
{
	for (unsigned n = 0; n < 3; ++n)
		for (unsigned i = 0; i < s->N[0]; ++i)
			for (unsigned j = 0; j < s->N[1]; ++j)
				for (unsigned k = 0; k < s->N[2]; ++k) { s->eng->SetVolt(n, i, j, k, 0); s->eng->SetCurr(n, i, j, k, 0); }
}

/* ---- readout through the reference's own engine interface / Processing classes */
double ref_voltage_integral(const ref_sim* s, const unsigned start[3], const unsigned stop[3])
{
	return s->eif->CalcVoltageIntegral(start, stop);
}
/* ProcessCurrent::CalcIntegral is protected and works on the snapped member box: build a ProcessCurrent,
   set its box directly and call it (Common/processcurrent.cpp:96-171) */
double ref_current_integral(const ref_sim* cs, const unsigned start[3], const unsigned stop[3], int norm_dir, const int start_inside[3], const int stop_inside[3])
{
	ref_sim* s = const_cast<ref_sim*>(cs);
	Engine_Interface_FDTD* eif;
#ifdef REF_WITH_CUDA
	if (dynamic_cast<Operator_CUDA*>(s->op)) eif = new Engine_Interface_CUDA_FDTD(dynamic_cast<Operator_CUDA*>(s->op));
	else
#endif
	if (Operator_sse* os = dynamic_cast<Operator_sse*>(s->op)) eif = new Engine_Interface_SSE_FDTD(os);
	else eif = new Engine_Interface_FDTD(s->op);
	ProcessCurrent pc(eif);   // owns eif
	for (int n = 0; n < 3; ++n) {
		pc.start[n] = start[n]; pc.stop[n] = stop[n];
		pc.m_start_inside[n] = start_inside[n] != 0; pc.m_stop_inside[n] = stop_inside[n] != 0;
	}
	pc.m_normDir = norm_dir;
	pc.m_Dimension = 2;
	return pc.CalcIntegral();
}
void ref_raw_field(const ref_sim* s, int is_H, const unsigned pos[3], double out[3])
{
	for (unsigned n = 0; n < 3; ++n) out[n] = is_H ? s->eif->GetRawDualField(n, pos, 0) : s->eif->GetRawField(n, pos, 0);
}
double ref_energy(const ref_sim* s) { return s->eif->CalcFastEnergy(); }
/* ProcessFields::CalcField Common/processfields.cpp:283-409 on the index box start..stop, no sub-sampling */
void ref_dump_field(const ref_sim* cs, int is_H, int interp, const unsigned start[3], const unsigned stop[3], float* out)
{
	ref_sim* s = const_cast<ref_sim*>(cs);
	Engine_Interface_FDTD* eif;
#ifdef REF_WITH_CUDA
	if (dynamic_cast<Operator_CUDA*>(s->op)) eif = new Engine_Interface_CUDA_FDTD(dynamic_cast<Operator_CUDA*>(s->op));
	else
#endif
	if (Operator_sse* os = dynamic_cast<Operator_sse*>(s->op)) eif = new Engine_Interface_SSE_FDTD(os);
	else eif = new Engine_Interface_FDTD(s->op);
	ProcessFieldsTD pf(eif);
	pf.SetDumpType(is_H ? ProcessFields::H_FIELD_DUMP : ProcessFields::E_FIELD_DUMP);
	eif->SetInterpolationType(interp == 0 ? Engine_Interface_Base::NO_INTERPOLATION
	                        : interp == 1 ? Engine_Interface_Base::NODE_INTERPOLATE : Engine_Interface_Base::CELL_INTERPOLATE);
	// the snapped box and sample positions InitProcess would set up (processfields.cpp:60-135) for no sub-sampling
	for (int n = 0; n < 3; ++n) {
		pf.start[n] = start[n]; pf.stop[n] = stop[n];
		pf.numLines[n] = stop[n] - start[n] + 1;
		delete[] pf.posLines[n]; delete[] pf.discLines[n];
		pf.posLines[n] = new unsigned int[pf.numLines[n]];
		pf.discLines[n] = new double[pf.numLines[n]];
		for (unsigned i = 0; i < pf.numLines[n]; ++i) { pf.posLines[n][i] = start[n] + i; pf.discLines[n][i] = s->op->GetDiscLine(n, start[n] + i, false); }
	}
	FDTD_FLOAT**** f = pf.CalcField();
	size_t p = 0;
	for (int n = 0; n < 3; ++n)
		for (unsigned k = 0; k < pf.numLines[2]; ++k)
			for (unsigned j = 0; j < pf.numLines[1]; ++j)
				for (unsigned i = 0; i < pf.numLines[0]; ++i) out[p++] = f[n][i][j][k];
	Delete_N_3DArray<FDTD_FLOAT>(f, pf.numLines);
}
double ref_edge_length(const ref_sim* s, int n, const unsigned pos[3], int dual) { return s->op->GetEdgeLength(n, pos, dual != 0); }
double ref_disc_line(const ref_sim* s, int n, unsigned pos, int dual) { return s->op->GetDiscLine(n, pos, dual != 0); }

/* ---- TFSF tables (Operator_Ext_TFSF) */
int ref_tfsf_on(const ref_sim* s) { return s->tfsf != NULL; }
unsigned ref_tfsf_max_delay(const ref_sim* s) { return s->tfsf ? s->tfsf->m_maxDelay : 0; }
void ref_tfsf_box(const ref_sim* s, unsigned start[3], unsigned stop[3], int active[6])
{
	for (int n = 0; n < 3; ++n) {
		start[n] = s->tfsf->m_Start[n]; stop[n] = s->tfsf->m_Stop[n];
		active[2 * n] = s->tfsf->m_ActiveDir[n][0]; active[2 * n + 1] = s->tfsf->m_ActiveDir[n][1];
	}
}
unsigned ref_tfsf_face(const ref_sim* s, int which, int n, int l, int c, unsigned* delay, float* delta, float* amp)
{
	Operator_Ext_TFSF* t = s->tfsf;
	int nP = (n + 1) % 3, nPP = (n + 2) % 3;
	unsigned numP = t->m_numLines[nP] * t->m_numLines[nPP];
	const unsigned* d = which ? t->m_CurrDelay[n][l][c] : t->m_VoltDelay[n][l][c];
	const FDTD_FLOAT* dd = which ? t->m_CurrDelayDelta[n][l][c] : t->m_VoltDelayDelta[n][l][c];
	const FDTD_FLOAT* a = which ? t->m_CurrAmp[n][l][c] : t->m_VoltAmp[n][l][c];
	if (!d) return 0;
	for (unsigned i = 0; i < numP; ++i) { delay[i] = d[i]; delta[i] = dd[i]; amp[i] = a[i]; }
	return numP;
}
/* ---- local absorbing sheets */
int ref_abc_count(const ref_sim* s) { return (int)s->abc.size(); }
void ref_abc_info(const ref_sim* s, int a, int* ny, int* type, int* positive, unsigned x0[3], unsigned x1[3])
{
	Operator_Ext_Absorbing_BC* o = s->abc[a];
	*ny = o->m_ny; *type = (int)o->m_ABCtype; *positive = o->m_normalSignPositive;
	for (int n = 0; n < 3; ++n) { x0[n] = o->m_sheetX0[n]; x1[n] = o->m_sheetX1[n]; }
}
void ref_abc_coeff(const ref_sim* s, int a, float* K1P, float* K1PP, float* K2P, float* K2PP)
{
	Operator_Ext_Absorbing_BC* o = s->abc[a];
	size_t p = 0;
	for (unsigned i = 0; i < o->m_numLines[0]; ++i)
		for (unsigned j = 0; j < o->m_numLines[1]; ++j, ++p) {
			K1P[p] = o->m_K1_nyP(i, j); K1PP[p] = o->m_K1_nyPP(i, j);
			// the K2 tables only exist with super-absorption (operator_ext_absorbing_bc.cpp: m_ABCtype == MUR_1ST_SA)
			bool sa = o->m_ABCtype == Operator_Ext_Absorbing_BC::MUR_1ST_SA;
			K2P[p] = sa ? o->m_K2_nyP(i, j) : 0; K2PP[p] = sa ? o->m_K2_nyPP(i, j) : 0;
		}
}

/* ---- the reference's Processing classes, run the way openEMS::RunFDTD does (openems.cpp:1393-1478) */
static Engine_Interface_FDTD* new_eif(ref_sim* s)
{
#ifdef REF_WITH_CUDA
	if (dynamic_cast<Operator_CUDA*>(s->op)) return new Engine_Interface_CUDA_FDTD(dynamic_cast<Operator_CUDA*>(s->op));
#endif
	if (Operator_sse* os = dynamic_cast<Operator_sse*>(s->op)) return new Engine_Interface_SSE_FDTD(os);
	return new Engine_Interface_FDTD(s->op);
}
/* kind: 0 voltage, 1 current, 2 E field probe, 3 H field probe (openEMS::SetupProcessing openems.cpp:560-640);
   start/stop in drawing units; the series is collected under `name` in the recorder-free way: the Processing
   object writes its ASCII file `name` into the current directory like the reference does */
int ref_add_probe(ref_sim* s, int kind, const char* name, const double start[3], const double stop[3], double weight, int norm_dir)
{
	ProcessIntegral* p = NULL;
	switch (kind) {
	case 0: p = new ProcessVoltage(new_eif(s)); break;
	case 1: {
		ProcessCurrent* c = NULL;
#ifdef REF_WITH_CUDA
		if (s->engine_kind == 4 && s->fast_processing) c = new ProcessCurrent_CUDA(new_eif(s));
#endif
		if (!c) c = new ProcessCurrent(new_eif(s));
		c->SetDualMesh(true); p = c; break; }
	case 2: p = new ProcessFieldProbe(new_eif(s), 0); break;
	case 3: { ProcessFieldProbe* f = new ProcessFieldProbe(new_eif(s), 1); f->SetDualMesh(true); p = f; break; }
	default: return -1;
	}
	if (kind == 1 || kind == 3) p->SetDualTime(true);
	p->SetProcessInterval(s->exc->GetNyquistNum() / 4);   // openems.cpp:568 with OverSampling 4 (openems.cpp:119)
	p->GetNormalDir(norm_dir);   // sic: the reference's setter is called GetNormalDir (processintegral.h:36, openems.cpp:572)
	p->SetName(name);
	double a[3] = { start[0], start[1], start[2] }, b[3] = { stop[0], stop[1], stop[2] };
	p->DefineStartStopCoord(a, b);
	p->SetWeight(weight);
	s->PA->AddProcessing(p);
	s->procs.push_back(NULL);   // owned by PA
	return (int)s->PA->GetNumberOfProcessings() - 1;
}
/* field dump box, dump_type 0 E 1 H, file_type 0 vtk 1 hdf5 (recorded in memory, see ref_glue.cpp), interp 0/1/2 */
int ref_add_dump(ref_sim* s, const char* name, const double start[3], const double stop[3], int dump_type, int file_type, int interp, unsigned interval)
{
	ProcessFieldsTD* p = NULL;
#ifdef REF_WITH_CUDA
	if (s->engine_kind == 4 && s->fast_processing) p = new ProcessFieldsTD_CUDA(new_eif(s));
#endif
	if (!p) p = new ProcessFieldsTD(new_eif(s));
	p->SetProcessInterval(interval ? interval : s->exc->GetNyquistNum() / 4);
	if (dump_type == 1) { p->SetDualTime(true); p->SetDualMesh(true); }   // openems.cpp:631-636
	p->SetDumpType((ProcessFields::DumpType)dump_type);
	p->SetDumpMode((Engine_Interface_Base::InterpolationType)interp);
	p->SetFileType(file_type ? ProcessFields::HDF5_FILETYPE : ProcessFields::VTK_FILETYPE);
	p->SetName(name);
	p->SetFileName(name);
	double a[3] = { start[0], start[1], start[2] }, b[3] = { stop[0], stop[1], stop[2] };
	p->DefineStartStopCoord(a, b);
	s->PA->AddProcessing(p);
	return (int)s->PA->GetNumberOfProcessings() - 1;
}
int ref_add_fd_dump(ref_sim* s, const char* name, const double start[3], const double stop[3], int dump_type, int interp, unsigned nfreq, const double* freqs)
{
	ProcessFieldsFD* p = NULL;
#ifdef REF_WITH_CUDA
	if (s->engine_kind == 4 && s->fast_processing) p = new ProcessFieldsFD_CUDA(new_eif(s));
#endif
	if (!p) p = new ProcessFieldsFD(new_eif(s));
	p->SetProcessInterval(s->exc->GetNyquistNum() / 4);
	if (dump_type == 1) { p->SetDualTime(true); p->SetDualMesh(true); }
	p->SetDumpType((ProcessFields::DumpType)dump_type);
	p->SetDumpMode((Engine_Interface_Base::InterpolationType)interp);
	p->SetFileType(ProcessFields::HDF5_FILETYPE);
	for (unsigned i = 0; i < nfreq; ++i) p->AddFrequency(freqs[i]);
	p->SetName(name);
	p->SetFileName(name);
	double a[3] = { start[0], start[1], start[2] }, b[3] = { stop[0], stop[1], stop[2] };
	p->DefineStartStopCoord(a, b);
	s->PA->AddProcessing(p);
	return (int)s->PA->GetNumberOfProcessings() - 1;
}
/* waveguide-port mode matching (openems.cpp:548-556): field_type 0 E / 1 H, mode functions of the two tangential
   directions in fparser syntax over x,y,z,rho,a,r,t */
int ref_add_mode_match(ref_sim* s, const char* name, const double start[3], const double stop[3], int field_type, const char* func_P, const char* func_PP, int ny)
{
	ProcessModeMatch* p = NULL;
#ifdef REF_WITH_CUDA
	if (s->engine_kind == 4 && s->fast_processing) p = new ProcessModeMatch_CUDA(new_eif(s));
#endif
	if (!p) p = new ProcessModeMatch(new_eif(s));
	p->SetFieldType(field_type);
	p->SetModeFunction((ny + 1) % 3, func_P);
	p->SetModeFunction((ny + 2) % 3, func_PP);
	if (field_type == 1) { p->SetDualTime(true); }
	p->SetProcessInterval(s->exc->GetNyquistNum() / 4);
	p->SetName(name);
	double a[3] = { start[0], start[1], start[2] }, b[3] = { stop[0], stop[1], stop[2] };
	p->DefineStartStopCoord(a, b);
	s->PA->AddProcessing(p);
	return (int)s->PA->GetNumberOfProcessings() - 1;
}

/* openEMS::RunFDTD main loop without the energy end criterion: IterateTS in steps of PA->Process() */
void ref_run(ref_sim* s, unsigned nr_ts)
{
	ftz_scope ftz;
	s->PA->InitAll();
	s->PA->PreProcess();
	int step = s->PA->Process();
	if ((step < 0) || (step > (int)nr_ts)) step = nr_ts;
	while (s->eng->GetNumberOfTimesteps() < nr_ts) {
		s->eng->IterateTS(step);
		step = s->PA->Process();
		int currTS = s->eng->GetNumberOfTimesteps();
		if ((step < 0) || (step > (int)(nr_ts - currTS))) step = nr_ts - currTS;
	}
	s->PA->FlushNext();
	s->PA->PostProcess();
}

/* ---- recorder access (what HDF5/VTK writers were asked to write) */
int ref_recorded_count(void) { return (int)ref_recorder_keys().size(); }
int ref_recorded_key(int i, char* out, int cap)
{
	std::vector<std::string> k = ref_recorder_keys();
	if (i < 0 || i >= (int)k.size()) return -1;
	strncpy(out, k[i].c_str(), cap - 1);
	out[cap - 1] = 0;
	return (int)k[i].size();
}
long ref_recorded_size(const char* key) { const ref_dataset* d = ref_recorder_get(key); return d ? (long)d->data.size() : -1; }
int ref_recorded_dims(const char* key, unsigned long* dims, int cap)
{
	const ref_dataset* d = ref_recorder_get(key);
	if (!d) return -1;
	for (int i = 0; i < cap && i < (int)d->dims.size(); ++i) dims[i] = d->dims[i];
	return (int)d->dims.size();
}
int ref_recorded_get(const char* key, double* out)
{
	const ref_dataset* d = ref_recorder_get(key);
	if (!d) return -1;
	std::copy(d->data.begin(), d->data.end(), out);
	return 0;
}
void ref_recorded_clear(void) { ref_recorder_clear(); }

/* number of de-duplicated f4 tuples of Operator_SSE_Compressed (a3), 0 for other operators */
unsigned ref_sse_unique(const ref_sim* s)
{
	const Operator_SSE_Compressed* c = dynamic_cast<const Operator_SSE_Compressed*>(s->op);
	return c ? (unsigned)c->f4_vv_Compressed[0].size() : 0;
}
const char* ref_version(void) { return GIT_VERSION; }

} // extern "C"
