/* engine_interface_cuda_fdtd.cpp -- see the header.  Goes to openEMS/FDTD/. */
#include "engine_interface_cuda_fdtd.h"

using namespace std;

const std::vector<double>& Engine_Interface_CUDA_FDTD::Values() const
{
	unsigned int ts = m_Eng_CUDA->GetNumberOfTimesteps();
	unsigned int n = 0;
	oems_cuda_num_probe_values(m_Eng_CUDA->GetHandle(), &n);
	if (ts!=m_cache_ts || m_cache.size()!=n)
	{
		m_cache.resize(n);
		if (oems_cuda_read_probes(m_Eng_CUDA->GetHandle(), m_cache.data()))
		{
			cerr << "Engine_Interface_CUDA_FDTD: " << oems_cuda_last_error(m_Eng_CUDA->GetHandle()) << endl;
			exit(2);
		}
		m_cache_ts = ts;
	}
	return m_cache;
}

int Engine_Interface_CUDA_FDTD::VoltageProbe(const unsigned int* start, const unsigned int* stop) const
{
	std::array<unsigned int,6> key = {start[0],start[1],start[2],stop[0],stop[1],stop[2]};
	std::map<std::array<unsigned int,6>, int>::const_iterator it = m_vprobes.find(key);
	if (it!=m_vprobes.end()) return it->second;
	int id=0;
	unsigned int slot=0;
	oems_cuda_num_probe_values(m_Eng_CUDA->GetHandle(), &slot);
	if (oems_cuda_add_probe_voltage(m_Eng_CUDA->GetHandle(), start, stop, &id)) return -1;
	m_vprobes[key] = (int)slot;
	m_cache_ts = (unsigned int)-1;
	return (int)slot;
}

int Engine_Interface_CUDA_FDTD::FieldProbe(int is_H, const unsigned int* pos) const
{
	std::array<unsigned int,4> key = {(unsigned int)is_H,pos[0],pos[1],pos[2]};
	std::map<std::array<unsigned int,4>, int>::const_iterator it = m_fprobes.find(key);
	if (it!=m_fprobes.end()) return it->second;
	int id=0;
	unsigned int slot=0;
	oems_cuda_num_probe_values(m_Eng_CUDA->GetHandle(), &slot);
	if (oems_cuda_add_probe_field(m_Eng_CUDA->GetHandle(), is_H, pos, &id)) return -1;
	m_fprobes[key] = (int)slot;
	m_cache_ts = (unsigned int)-1;
	return (int)slot;
}

double Engine_Interface_CUDA_FDTD::CalcVoltageIntegral(const unsigned int* start, const unsigned int* stop) const
{
	if (m_Eng_CUDA==NULL) return Engine_Interface_FDTD::CalcVoltageIntegral(start,stop);
	if (((start[0]!=stop[0]) + (start[1]!=stop[1]) + (start[2]!=stop[2]))!=1)
	{
		cerr << "Engine_Interface_CUDA_FDTD::CalcVoltageIntegral: Error, only a 1D/line integration is allowed" << endl;
		return 0;
	}
	int slot = VoltageProbe(start,stop);
	if (slot<0) return Engine_Interface_FDTD::CalcVoltageIntegral(start,stop);
	return Values().at(slot);
}

double Engine_Interface_CUDA_FDTD::GetRawField(unsigned int n, const unsigned int* pos, int type) const
{
	if (m_Eng_CUDA==NULL || type!=0) return Engine_Interface_FDTD::GetRawField(n,pos,type);
	int slot = FieldProbe(0,pos);
	if (slot<0) return Engine_Interface_FDTD::GetRawField(n,pos,type);
	double value = Values().at(slot+n);
	double delta = m_Op->GetEdgeLength(n,pos);
	if (delta) return value/delta;
	return 0.0;
}

double Engine_Interface_CUDA_FDTD::GetRawDualField(unsigned int n, const unsigned int* pos, int type) const
{
	if (m_Eng_CUDA==NULL || type!=0) return Engine_Interface_FDTD::GetRawDualField(n,pos,type);
	int slot = FieldProbe(1,pos);
	if (slot<0) return Engine_Interface_FDTD::GetRawDualField(n,pos,type);
	double value = Values().at(slot+n);
	double delta = m_Op->GetEdgeLength(n,pos,true);
	if (delta) return value/delta;
	return 0.0;
}

double Engine_Interface_CUDA_FDTD::CalcFastEnergy() const
{
	if (m_Eng_CUDA==NULL) return Engine_Interface_FDTD::CalcFastEnergy();
	double e=0;
	if (oems_cuda_energy(m_Eng_CUDA->GetHandle(), &e)) return Engine_Interface_FDTD::CalcFastEnergy();
	return e;
}
