/* processing_cuda.cpp -- see processing_cuda.h.  Goes to openEMS/Common/processing_cuda.cpp. */
#include "processing_cuda.h"
#include "FDTD/engine_interface_fdtd.h"
#include "Common/operator_base.h"
#include "tools/hdf5_file_writer.h"
#include "tools/vtk_file_writer.h"
#include "tools/array_ops.h"

#include <iomanip>
#include <sstream>
#include <complex>
#include <cmath>

using namespace std;

static Engine_CUDA* cuda_engine_of(Engine_Interface_Base* eng_if)
{
	Engine_Interface_FDTD* ei = dynamic_cast<Engine_Interface_FDTD*>(eng_if);
	if (ei==NULL) return NULL;
	return dynamic_cast<Engine_CUDA*>(const_cast<Engine*>(ei->GetFDTDEngine()));
}

static void fail(Engine_CUDA* eng, const char* what)
{
	cerr << what << ": " << oems_cuda_last_error(eng->GetHandle()) << endl;
	exit(2);
}

//! Operator::GetEdgeLength(n,pos,dual) along direction n (Cartesian mesh: independent of the other two indices)
static void edge_length_tables(const Operator_Base* op, std::vector<double> el[3], std::vector<double> del[3])
{
	for (int n=0; n<3; ++n)
	{
		unsigned int cnt = op->GetNumberOfLines(n, true);
		el[n].resize(cnt); del[n].resize(cnt);
		for (unsigned int i=0; i<cnt; ++i)
		{
			unsigned int pos[3] = {0,0,0};
			pos[n] = i;
			el[n][i] = op->GetEdgeLength(n,pos,false);
			del[n][i] = op->GetEdgeLength(n,pos,true);
		}
	}
}

// ------------------------------------------------------------------------------ ProcessCurrent_CUDA
double ProcessCurrent_CUDA::CalcIntegral()
{
	Engine_CUDA* eng = cuda_engine_of(m_Eng_Interface);
	if (eng==NULL) return ProcessCurrent::CalcIntegral();
	if (m_probe<0)
	{
		int si[3], ei[3];
		for (int n=0; n<3; ++n) { si[n] = m_start_inside[n]; ei[n] = m_stop_inside[n]; }
		oems_cuda_num_probe_values(eng->GetHandle(), &m_slot);
		if (oems_cuda_add_probe_current(eng->GetHandle(), start, stop, m_normDir, si, ei, &m_probe))
			fail(eng, "ProcessCurrent_CUDA");
	}
	unsigned int n=0;
	oems_cuda_num_probe_values(eng->GetHandle(), &n);
	m_values.resize(n);
	if (oems_cuda_read_probes(eng->GetHandle(), m_values.data())) fail(eng, "ProcessCurrent_CUDA");
	return m_values.at(m_slot);
}

// ------------------------------------------------------------------------------ dump box shared by TD / FD
DeviceDumpBox::~DeviceDumpBox()
{
	if (pinned) oems_cuda_host_free(pinned);
}

bool DeviceDumpBox::Setup(Engine_Interface_Base* eng_if, const Operator_Base* op, int dump_type, const unsigned int numLines[3], unsigned int* const posLines[3])
{
	eng = cuda_engine_of(eng_if);
	if (eng==NULL) return false;
	if ((dump_type!=ProcessFields::E_FIELD_DUMP) && (dump_type!=ProcessFields::H_FIELD_DUMP)) return false; // J, rotH, D, B: stock path
	int interp = 0;
	switch (eng_if->GetInterpolationType())
	{
	case Engine_Interface_Base::NO_INTERPOLATION: interp = 0; break;
	case Engine_Interface_Base::NODE_INTERPOLATE: interp = 1; break;
	case Engine_Interface_Base::CELL_INTERPOLATE: interp = 2; break;
	default: return false;
	}
	std::vector<double> el[3], del[3];
	edge_length_tables(op, el, del);
	const double* pel[3] = {el[0].data(), el[1].data(), el[2].data()};
	const double* pdel[3] = {del[0].data(), del[1].data(), del[2].data()};
	if (oems_cuda_add_dump(eng->GetHandle(), dump_type==ProcessFields::H_FIELD_DUMP, interp, numLines[0], numLines[1], numLines[2],
		posLines[0], posLines[1], posLines[2], pel, pdel, &dump_id))
		fail(eng, "DeviceDumpBox::Setup");
	count = (size_t)3*numLines[0]*numLines[1]*numLines[2];
	void* p = NULL;
	if (oems_cuda_host_alloc(count*sizeof(float), &p)) fail(eng, "DeviceDumpBox::Setup (pinned buffer)");
	pinned = (float*)p;
	return true;
}

// ------------------------------------------------------------------------------ ProcessFieldsTD_CUDA
ProcessFieldsTD_CUDA::~ProcessFieldsTD_CUDA()
{
}

void ProcessFieldsTD_CUDA::InitProcess()
{
	ProcessFieldsTD::InitProcess();
	if (Enabled==false) return;
	m_box.Setup(m_Eng_Interface, Op, m_DumpType, numLines, posLines);
}

bool ProcessFieldsTD_CUDA::WriteOut(unsigned int ts, float time_value)
{
	if (oems_cuda_wait(m_box.eng->GetHandle(), m_box.ticket)) fail(m_box.eng, "ProcessFieldsTD_CUDA");
	bool success = true;
	if (m_fileType==VTK_FILETYPE)
	{
		// the VTK writer takes N-I-J-K arrays: un-flatten {3,nz,ny,nx}
		FDTD_FLOAT**** field = Create_N_3DArray<FDTD_FLOAT>(numLines);
		size_t p=0;
		for (int n=0; n<3; ++n)
			for (unsigned int k=0; k<numLines[2]; ++k)
				for (unsigned int j=0; j<numLines[1]; ++j)
					for (unsigned int i=0; i<numLines[0]; ++i)
						field[n][i][j][k] = m_box.pinned[p++];
		m_Vtk_Dump_File->SetTimestep(ts);
		m_Vtk_Dump_File->ClearAllFields();
		m_Vtk_Dump_File->AddVectorField(GetFieldNameByType(m_DumpType),field);
		success &= m_Vtk_Dump_File->Write();
		Delete_N_3DArray<FDTD_FLOAT>(field,numLines);
	}
	else if (m_fileType==HDF5_FILETYPE)
	{
		// the device block already has the layout HDF5_File_Writer::WriteVectorField produces
		// (tools/hdf5_file_writer.cpp:286-302): hand it to the public flat writer
		stringstream ss;
		ss << std::setw( pad_length ) << std::setfill( '0' ) << ts;
		size_t n_size[4]={3,numLines[2],numLines[1],numLines[0]};
		success &= m_HDF5_Dump_File->WriteData(ss.str(), m_box.pinned, 4, n_size);
		float time[1] = {time_value};
		success &= m_HDF5_Dump_File->WriteAtrribute("/FieldData/TD/"+ss.str(),"time",time,1);
	}
	else
		success = false;
	return success;
}

int ProcessFieldsTD_CUDA::Process()
{
	if (m_box.eng==NULL) return ProcessFieldsTD::Process();
	if (Enabled==false) return -1;
	if (CheckTimestep()==false) return GetNextInterval();

	bool success = true;
	// write the previous dump (its copy ran while the engine was stepping), then launch this one
	if (m_pending)
		success &= WriteOut(m_pending_ts, m_pending_time);
	if (oems_cuda_read_dump_async(m_box.eng->GetHandle(), m_box.dump_id, m_box.pinned, &m_box.ticket))
		fail(m_box.eng, "ProcessFieldsTD_CUDA::Process");
	m_pending = true;
	m_pending_ts = m_Eng_Interface->GetNumberOfTimesteps();
	m_pending_time = (float)m_Eng_Interface->GetTime(m_dualTime);

	if (success==false)
	{
		SetEnable(false);
		cerr << "ProcessFieldsTD_CUDA::Process: can't dump to file... disabled! " << endl;
	}
	return GetNextInterval();
}

void ProcessFieldsTD_CUDA::PostProcess()
{
	if (m_pending && m_box.eng)
	{
		WriteOut(m_pending_ts, m_pending_time);
		m_pending = false;
	}
	ProcessFieldsTD::PostProcess();
}

// ------------------------------------------------------------------------------ ProcessFieldsFD_CUDA
void ProcessFieldsFD_CUDA::InitProcess()
{
	ProcessFieldsFD::InitProcess();   // also allocates the host accumulators m_FD_Fields used by DumpFDData
	if (Enabled==false) return;
	if (!m_box.Setup(m_Eng_Interface, Op, m_DumpType, numLines, posLines)) return;
	if (oems_cuda_add_fd_dump(m_box.eng->GetHandle(), m_box.dump_id, m_FD_Samples.size(), &m_fd_id))
		fail(m_box.eng, "ProcessFieldsFD_CUDA::InitProcess");
}

int ProcessFieldsFD_CUDA::Process()
{
	if (m_box.eng==NULL || m_fd_id<0) return ProcessFieldsFD::Process();
	if (Enabled==false) return -1;
	if (CheckTimestep()==false) return GetNextInterval();

	if ((m_FD_Interval==0) || (m_Eng_Interface->GetNumberOfTimesteps()%m_FD_Interval!=0))
		return GetNextInterval();

	// the weights exactly as processfields_fd.cpp:84-86 computes them (complex<float> arithmetic on the host)
	double T = m_Eng_Interface->GetTime(m_dualTime);
	std::vector<float> w(2*m_FD_Samples.size());
	for (size_t n = 0; n<m_FD_Samples.size(); ++n)
	{
		std::complex<float> exp_jwt_2_dt = std::exp( (std::complex<float>)(-2.0 * _I * M_PI * m_FD_Samples.at(n) * T) );
		exp_jwt_2_dt *= 2; // *2 for single-sided spectrum
		exp_jwt_2_dt *= Op->GetTimestep() * m_FD_Interval; // multiply with timestep-interval
		w[2*n] = exp_jwt_2_dt.real();
		w[2*n+1] = exp_jwt_2_dt.imag();
	}
	if (oems_cuda_fd_accumulate(m_box.eng->GetHandle(), m_fd_id, w.data()))
		fail(m_box.eng, "ProcessFieldsFD_CUDA::Process");
	++m_FD_SampleCount;
	return GetNextInterval();
}

void ProcessFieldsFD_CUDA::DumpFDData()
{
	if (m_box.eng && m_fd_id>=0)
	{
		// one D2H of the device accumulators into the host arrays the stock writer code reads
		const size_t per = (size_t)3*numLines[0]*numLines[1]*numLines[2];
		std::vector<float> buf(2*per*m_FD_Samples.size());
		unsigned int samples=0;
		if (oems_cuda_read_fd(m_box.eng->GetHandle(), m_fd_id, buf.data(), &samples))
			fail(m_box.eng, "ProcessFieldsFD_CUDA::DumpFDData");
		for (size_t f = 0; f<m_FD_Samples.size(); ++f)
		{
			std::complex<float>**** field_fd = m_FD_Fields.at(f);
			size_t p = 2*per*f;
			for (int n=0; n<3; ++n)
				for (unsigned int k=0; k<numLines[2]; ++k)
					for (unsigned int j=0; j<numLines[1]; ++j)
						for (unsigned int i=0; i<numLines[0]; ++i, p+=2)
							field_fd[n][i][j][k] = std::complex<float>(buf[p], buf[p+1]);
		}
	}
	ProcessFieldsFD::DumpFDData();
}

// ------------------------------------------------------------------------------ ProcessModeMatch_CUDA
void ProcessModeMatch_CUDA::InitProcess()
{
	ProcessModeMatch::InitProcess();   // sorts/shrinks the box, parses and normalises the mode template
	if (!Enabled) return;
	m_eng = cuda_engine_of(m_Eng_Interface);
	if (m_eng==NULL) return;
	const bool dualMesh = m_ModeFieldType==1;
	const int nP = (m_ny+1)%3, nPP = (m_ny+2)%3;
	std::vector<double> d0((size_t)m_numLines[0]*m_numLines[1]), d1(d0.size()), area(d0.size());
	unsigned int pos[3] = {0,0,0};
	pos[m_ny] = start[m_ny];
	for (unsigned int posP = 0; posP<m_numLines[0]; ++posP)
	{
		pos[nP] = start[nP] + posP;
		for (unsigned int posPP = 0; posPP<m_numLines[1]; ++posPP)
		{
			pos[nPP] = start[nPP] + posPP;
			const size_t q = (size_t)posP*m_numLines[1]+posPP;
			d0[q] = m_ModeDist[0][posP][posPP];
			d1[q] = m_ModeDist[1][posP][posPP];
			area[q] = Op->GetNodeArea(m_ny,pos,dualMesh);
		}
	}
	std::vector<double> el[3], del[3];
	edge_length_tables(Op, el, del);
	const double* pel[3] = {el[0].data(), el[1].data(), el[2].data()};
	const double* pdel[3] = {del[0].data(), del[1].data(), del[2].data()};
	if (oems_cuda_add_mode_match(m_eng->GetHandle(), m_ModeFieldType, m_ny, start, stop, d0.data(), d1.data(), area.data(), pel, pdel, &m_id))
		fail(m_eng, "ProcessModeMatch_CUDA::InitProcess");
}

double* ProcessModeMatch_CUDA::CalcMultipleIntegrals()
{
	if (m_eng==NULL || m_id<0) return ProcessModeMatch::CalcMultipleIntegrals();
	double out[2] = {0,0};
	if (oems_cuda_read_mode_match(m_eng->GetHandle(), m_id, out))
		fail(m_eng, "ProcessModeMatch_CUDA::CalcMultipleIntegrals");
	m_Results[0] = out[0];
	m_Results[1] = out[1];
	return m_Results;
}
