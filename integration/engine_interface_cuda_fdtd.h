/*
 * engine_interface_cuda_fdtd.h -- Engine_Interface_FDTD served from the device engine.
 * Goes to openEMS/FDTD/engine_interface_cuda_fdtd.h.  Compiled inside openEMS.
 *
 * Processing objects keep calling the same virtuals; line integrals and point fields are
 * registered as device probes on first use and then answered by one batched device
 * reduction per Process() round instead of one virtual GetVolt per edge.
 */
#ifndef ENGINE_INTERFACE_CUDA_FDTD_H
#define ENGINE_INTERFACE_CUDA_FDTD_H

#include "engine_interface_fdtd.h"
#include "engine_cuda.h"
#include <map>
#include <vector>
#include <array>

class Engine_Interface_CUDA_FDTD : public Engine_Interface_FDTD
{
public:
	//! Engine_Interface_FDTD's constructor takes the engine from op->GetEngine() (engine_interface_fdtd.cpp:31):
	//! interfaces are created after Operator::CreateEngine (openems.cpp:1316-1330)
	Engine_Interface_CUDA_FDTD(Operator* op) : Engine_Interface_FDTD(op), m_Eng_CUDA(dynamic_cast<Engine_CUDA*>(m_Eng)),
		m_slots(0), m_cache_ts((unsigned int)-1) {}
	virtual ~Engine_Interface_CUDA_FDTD() {}

	virtual std::string GetInterfaceString() const {return std::string("B200 CUDA FDTD engine interface");}

	virtual void SetFDTDEngine(Engine* eng) {Engine_Interface_FDTD::SetFDTDEngine(eng); m_Eng_CUDA = dynamic_cast<Engine_CUDA*>(eng);}

	virtual double CalcVoltageIntegral(const unsigned int* start, const unsigned int* stop) const;
	virtual double CalcFastEnergy() const;

protected:
	//! point fields: GetRawField/GetRawDualField go through the probe cache
	virtual double GetRawField(unsigned int n, const unsigned int* pos, int type) const;
	virtual double GetRawDualField(unsigned int n, const unsigned int* pos, int type) const;

	const std::vector<double>& Values() const;
	int VoltageProbe(const unsigned int* start, const unsigned int* stop) const;
	int FieldProbe(int is_H, const unsigned int* pos) const;

	Engine_CUDA* m_Eng_CUDA;
	mutable std::map<std::array<unsigned int,6>, int> m_vprobes;   // key -> first value slot
	mutable std::map<std::array<unsigned int,4>, int> m_fprobes;
	mutable unsigned int m_slots;
	mutable std::vector<double> m_cache;
	mutable unsigned int m_cache_ts;
};

#endif
