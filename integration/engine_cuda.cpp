/*
 * engine_cuda.cpp -- see engine_cuda.h.  Goes to openEMS/FDTD/engine_cuda.cpp.
 *
 * The stock Engine_Ext_* objects cannot be reused: they go through ENG_DISPATCH, which throws
 * for any engine type other than BASIC/SSE (extensions/engine_extension_dispatcher.h:58-72).
 * InitExtensions therefore reads the OPERATOR extensions' tables (one `friend class
 * Engine_CUDA;` line per operator_ext header) and hands them to the library.
 */
#include "engine_cuda.h"
#include <thread>
#include <algorithm>
#include "operator_cuda.h"
#include "extensions/operator_ext_excitation.h"
#include "extensions/operator_ext_upml.h"
#include "extensions/operator_ext_mur_abc.h"
#include "extensions/operator_ext_lorentzmaterial.h"
#include "extensions/operator_ext_conductingsheet.h"
#include "extensions/operator_ext_lumpedRLC.h"
#include "extensions/operator_ext_steadystate.h"
#include "extensions/engine_ext_steadystate.h"
#include "extensions/operator_ext_tfsf.h"
#include "extensions/operator_ext_absorbing_bc.h"
#include "extensions/engine_extension.h"
#include "excitation.h"

#include <cstdlib>
#include <vector>

using namespace std;

Engine_CUDA* Engine_CUDA::New(const Operator_CUDA* op)
{
	cout << "Create FDTD engine (B200 CUDA)" << endl;
	Engine_CUDA* e = new Engine_CUDA(op);
	e->Init();
	return e;
}

Engine_CUDA::Engine_CUDA(const Operator_CUDA* op) : Engine(op), m_Op_CUDA(op), m_h(NULL), m_SSD(NULL)
{
	m_type = UNKNOWN; // neither BASIC nor SSE: stock extensions must not dispatch on this engine
}

Engine_CUDA::~Engine_CUDA()
{
	Reset();
}

void Engine_CUDA::Check(int rc, const char* what) const
{
	if (rc==0) return;
	cerr << "Engine_CUDA::" << what << ": " << oems_cuda_last_error(m_h) << endl;
	exit(2); // the reference's convention for unrecoverable engine errors (openems.cpp:803-807)
}

void Engine_CUDA::Reset()
{
	if (m_h) oems_cuda_destroy(m_h);
	m_h = NULL;
	Engine::Reset();
}

void Engine_CUDA::Init()
{
	numTS = 0;
	// no host field arrays: volt_ptr / curr_ptr stay NULL (all accessors are overridden)
	Check( oems_cuda_create(numLines[0], numLines[1], numLines[2], m_Op_CUDA->GetDevice(), &m_h), "Init" );

	// dense coefficients in ArrayNIJK order, as the C ABI wants them.  The operator keeps them in the lane-interleaved
	// f4vector layout of Operator_sse (z -> (z % numVectors, lane z / numVectors), operator_sse.h:35-57), so this is a
	// gather through the accessors, done once and spread over the host threads (one x range each)
	const size_t N = (size_t)numLines[0]*numLines[1]*numLines[2];
	std::vector<FDTD_FLOAT> vv(3*N), vi(3*N), ii(3*N), iv(3*N);
	{
		unsigned int nt = std::max(1u, std::min(std::thread::hardware_concurrency(), numLines[0]));
		std::vector<std::thread> pool;
		for (unsigned int t=0; t<nt; ++t)
			pool.emplace_back([&, t]()
			{
				const unsigned int x0 = (unsigned int)((size_t)numLines[0]*t/nt), x1 = (unsigned int)((size_t)numLines[0]*(t+1)/nt);
				for (unsigned int n=0; n<3; ++n)
					for (unsigned int x=x0; x<x1; ++x)
						for (unsigned int y=0; y<numLines[1]; ++y)
						{
							size_t p = (((size_t)n*numLines[0] + x)*numLines[1] + y)*numLines[2];
							for (unsigned int z=0; z<numLines[2]; ++z, ++p)
							{
								vv[p] = Op->GetVV(n,x,y,z); vi[p] = Op->GetVI(n,x,y,z);
								ii[p] = Op->GetII(n,x,y,z); iv[p] = Op->GetIV(n,x,y,z);
							}
						}
			});
		for (std::thread& th : pool) th.join();
	}
	Check( oems_cuda_set_operator_dense(m_h, vv.data(), vi.data(), ii.data(), iv.data()), "set_operator" );

	InitExtensions();
	Check( oems_cuda_finalize(m_h), "finalize" );
}

void Engine_CUDA::InitExtensions()
{
	Excitation* exc = Op->GetExcitationSignal();
	unsigned int period_ts = 0;
	if (exc->GetSignalPeriod()>0)
		period_ts = (unsigned int)int(exc->GetSignalPeriod()/exc->GetTimestep());
	Check( oems_cuda_set_signal(m_h, exc->GetVoltageSignal(), exc->GetCurrentSignal(), exc->GetLength(), period_ts), "set_signal" );

	for (size_t n=0; n<Op->GetNumberOfExtentions(); ++n)
	{
		Operator_Extension* op_ext = Op->GetExtension(n);

		if (Operator_Ext_Excitation* e = dynamic_cast<Operator_Ext_Excitation*>(op_ext))
		{
			for (int w=0; w<2; ++w)
			{
				unsigned int cnt = w ? e->Curr_Count : e->Volt_Count;
				if (cnt==0) continue;
				unsigned int** idx = w ? e->Curr_index : e->Volt_index;
				unsigned short* d16 = w ? e->Curr_dir : e->Volt_dir;
				std::vector<unsigned int> idx3(3*(size_t)cnt), dir(cnt);
				for (unsigned int i=0; i<cnt; ++i)
				{
					for (int a=0; a<3; ++a) idx3[(size_t)a*cnt+i] = idx[a][i];
					dir[i] = d16[i];
				}
				Check( oems_cuda_add_excitation(m_h, w, cnt, idx3.data(), dir.data(),
					w ? e->Curr_amp : e->Volt_amp, w ? e->Curr_delay : e->Volt_delay), "add_excitation" );
			}
			continue;
		}
		if (Operator_Ext_UPML* u = dynamic_cast<Operator_Ext_UPML*>(op_ext))
		{
			Check( oems_cuda_add_upml(m_h, u->m_StartPos, u->m_numLines, u->vv.data(), u->vvfn.data(), u->vvfo.data(),
				u->ii.data(), u->iifn.data(), u->iifo.data()), "add_upml" );
			continue;
		}
		if (Operator_Ext_Mur_ABC* m = dynamic_cast<Operator_Ext_Mur_ABC*>(op_ext))
		{
			// delayed start, Engine_Ext_Mur_ABC ctor engine_ext_mur_abc.cpp:44-60
			int maxDelay=-1;
			Operator_Ext_Excitation* Exc_ext = Op->GetExcitationExtension();
			for (unsigned int i=0; i<Exc_ext->GetVoltCount(); ++i)
				if ( ((Exc_ext->Volt_dir[i]==m->m_nyP) || (Exc_ext->Volt_dir[i]==m->m_nyPP)) && (Exc_ext->Volt_index[m->m_ny][i]==m->m_LineNr) )
					if ((int)Exc_ext->Volt_delay[i]>maxDelay) maxDelay = (int)Exc_ext->Volt_delay[i];
			unsigned int start_ts = (maxDelay>=0) ? maxDelay + exc->GetLength() + 10 : 0;
			Check( oems_cuda_add_mur(m_h, m->m_ny, m->m_LineNr, m->m_LineNr_Shift, m->m_numLines,
				m->m_Mur_Coeff_nyP.data(), m->m_Mur_Coeff_nyPP.data(), start_ts), "add_mur" );
			continue;
		}
		// Operator_Ext_ConductingSheet derives from Operator_Ext_LorentzMaterial: same tables
		if (Operator_Ext_LorentzMaterial* l = dynamic_cast<Operator_Ext_LorentzMaterial*>(op_ext))
		{
			for (int o=0; o<l->m_Order; ++o)
			{
				unsigned int cnt = l->m_LM_Count.at(o);
				std::vector<unsigned int> pos3(3*(size_t)cnt);
				std::vector<FDTD_FLOAT> c[6];
				FDTD_FLOAT*** src[6] = {l->v_int_ADE, l->v_ext_ADE, l->v_Lor_ADE, l->i_int_ADE, l->i_ext_ADE, l->i_Lor_ADE};
				bool on[6] = {l->m_volt_ADE_On[o], l->m_volt_ADE_On[o], l->m_volt_Lor_ADE_On[o],
				              l->m_curr_ADE_On[o], l->m_curr_ADE_On[o], l->m_curr_Lor_ADE_On[o]};
				for (int a=0; a<3; ++a)
					for (unsigned int i=0; i<cnt; ++i) pos3[(size_t)a*cnt+i] = l->m_LM_pos[o][a][i];
				for (int w=0; w<6; ++w)
					if (on[w])
					{
						c[w].resize(3*(size_t)cnt);
						for (int a=0; a<3; ++a)
							for (unsigned int i=0; i<cnt; ++i) c[w][(size_t)a*cnt+i] = src[w][o][a][i];
					}
				Check( oems_cuda_add_lorentz(m_h, cnt, pos3.data(), on[0]?c[0].data():NULL, on[1]?c[1].data():NULL, on[2]?c[2].data():NULL,
					on[3]?c[3].data():NULL, on[4]?c[4].data():NULL, on[5]?c[5].data():NULL), "add_lorentz" );
			}
			continue;
		}
		if (Operator_Ext_LumpedRLC* r = dynamic_cast<Operator_Ext_LumpedRLC*>(op_ext))
		{
			unsigned int cnt = r->RLC_count;
			if (cnt==0) continue;
			std::vector<unsigned int> pos3(3*(size_t)cnt);
			for (int a=0; a<3; ++a)
				for (unsigned int i=0; i<cnt; ++i) pos3[(size_t)a*cnt+i] = r->v_RLC_pos[a][i];
			Check( oems_cuda_add_rlc(m_h, cnt, r->v_RLC_dir, pos3.data(), r->v_RLC_ilv, r->v_RLC_i2v, r->v_RLC_vvd, r->v_RLC_vv2,
				r->v_RLC_vj1, r->v_RLC_vj2, r->v_RLC_ib0, r->v_RLC_b1, r->v_RLC_b2), "add_rlc" );
			continue;
		}
		if (Operator_Ext_SteadyState* ss = dynamic_cast<Operator_Ext_SteadyState*>(op_ext))
		{
			// recorded and evaluated on the device (oems_cuda_add_steadystate); the stock engine extension is
			// still created because the driver dereferences it without a NULL check (openems.cpp:1318-1322) and
			// reads GetLastDiff() from it (:1465) -- IterateTS stores the device result there, its own
			// Apply2Voltages (one GetVolt per probe and timestep) is never called.
			const unsigned int cnt = ss->m_E_probe_dir.size();
			std::vector<unsigned int> pos3(3*(size_t)cnt);
			for (int a=0; a<3; ++a)
				for (unsigned int i=0; i<cnt; ++i) pos3[(size_t)a*cnt+i] = ss->m_E_probe_pos[a].at(i);
			Check( oems_cuda_add_steadystate(m_h, ss->m_TS_period, cnt, pos3.data(), ss->m_E_probe_dir.data()), "add_steadystate" );
			m_SSD = dynamic_cast<Engine_Ext_SteadyState*>(op_ext->CreateEngineExtention());
			if (m_SSD)
				m_SSD->SetEngine(this);
			continue;
		}
		if (Operator_Ext_TFSF* t = dynamic_cast<Operator_Ext_TFSF*>(op_ext))
		{
			if (!t->IsActive()) continue; // no plane-wave excitation in this setup (operator_ext_tfsf.cpp:103-108)
			int active[6];
			const unsigned int* vd[12]; const FDTD_FLOAT* vdd[12]; const FDTD_FLOAT* va[12];
			const unsigned int* cd[12]; const FDTD_FLOAT* cdd[12]; const FDTD_FLOAT* ca[12];
			for (int n=0; n<3; ++n)
				for (int l=0; l<2; ++l)
				{
					active[2*n+l] = t->m_ActiveDir[n][l];
					for (int c=0; c<2; ++c)
					{
						const int q = (n*2+l)*2+c;
						vd[q] = t->m_VoltDelay[n][l][c]; vdd[q] = t->m_VoltDelayDelta[n][l][c]; va[q] = t->m_VoltAmp[n][l][c];
						cd[q] = t->m_CurrDelay[n][l][c]; cdd[q] = t->m_CurrDelayDelta[n][l][c]; ca[q] = t->m_CurrAmp[n][l][c];
					}
				}
			Check( oems_cuda_set_tfsf(m_h, t->m_Start, t->m_Stop, active, vd, vdd, va, cd, cdd, ca), "set_tfsf" );
			continue;
		}
		if (Operator_Ext_Absorbing_BC* a = dynamic_cast<Operator_Ext_Absorbing_BC*>(op_ext))
		{
			const bool sa = a->m_ABCtype==Operator_Ext_Absorbing_BC::MUR_1ST_SA;
			// ArrayIJ<FDTD_FLOAT> is one contiguous [i][j] block (tools/arraylib/array_ij.h)
			Check( oems_cuda_add_absorbing_sheet(m_h, a->m_ny, a->m_sheetX0, a->m_sheetX1, a->m_normalSignPositive, (int)a->m_ABCtype,
				&a->m_K1_nyP(0,0), &a->m_K1_nyPP(0,0), sa ? &a->m_K2_nyP(0,0) : NULL, sa ? &a->m_K2_nyPP(0,0) : NULL), "add_absorbing_sheet" );
			continue;
		}
		cerr << "Engine_CUDA::InitExtensions: extension \"" << op_ext->GetExtensionName()
		     << "\" has no device implementation (cylinder extensions: see DESIGN.md), aborting" << endl;
		exit(2);
	}
}

bool Engine_CUDA::IterateTS(unsigned int iterTS)
{
	Check( oems_cuda_iterate(m_h, iterTS), "IterateTS" );
	numTS += iterTS;
	if (m_SSD)
	{
		// what Engine_Ext_SteadyState::Apply2Voltages would have left behind (engine_ext_steadystate.cpp:50-107)
		double diff = 1; unsigned int checks = 0;
		Check( oems_cuda_steadystate_check(m_h, &diff, &checks), "steadystate_check" );
		m_SSD->m_last_max_diff = diff;
	}
	return true;
}

unsigned int Engine_CUDA::GetNumberOfTimesteps()
{
	return numTS;
}

FDTD_FLOAT Engine_CUDA::GetVolt(unsigned int n, unsigned int x, unsigned int y, unsigned int z) const
{
	float v=0;
	Check( oems_cuda_get_field(m_h, 0, n, x, y, z, &v), "GetVolt" );
	return v;
}
FDTD_FLOAT Engine_CUDA::GetCurr(unsigned int n, unsigned int x, unsigned int y, unsigned int z) const
{
	float v=0;
	Check( oems_cuda_get_field(m_h, 1, n, x, y, z, &v), "GetCurr" );
	return v;
}
void Engine_CUDA::SetVolt(unsigned int n, unsigned int x, unsigned int y, unsigned int z, FDTD_FLOAT value)
{
	Check( oems_cuda_set_field(m_h, 0, n, x, y, z, value), "SetVolt" );
}
void Engine_CUDA::SetCurr(unsigned int n, unsigned int x, unsigned int y, unsigned int z, FDTD_FLOAT value)
{
	Check( oems_cuda_set_field(m_h, 1, n, x, y, z, value), "SetCurr" );
}
