/*
 * processing_cuda.h -- device-served variants of the four Processing classes whose stock implementation walks
 * the mesh through per-cell virtual accessors.  Goes to openEMS/Common/processing_cuda.h; compiled inside
 * openEMS.  In this repository it is compiled against the reference's unmodified headers and RUN by the
 * oracle/_ref harness (oracle/Makefile.ref target libopenems_ref_cuda.so, tests/test_gpu_reference_integration.py).
 *
 * The stock classes work unchanged on Engine_CUDA (through Engine_Interface_CUDA_FDTD and the slow per-cell
 * accessors); these subclasses replace only the inner gather:
 *   ProcessCurrent_CUDA     CalcIntegral()            Common/processcurrent.cpp:96-171    -> one device probe
 *   ProcessFieldsTD_CUDA    Process()                 Common/processfields_td.cpp:50-91   -> k_dump, async D2H
 *   ProcessFieldsFD_CUDA    Process()/DumpFDData()    Common/processfields_fd.cpp:72-230  -> device running DFT
 *   ProcessModeMatch_CUDA   CalcMultipleIntegrals()   Common/processmodematch.cpp:222-266 -> k_mode_match
 * Hook: openEMS::SetupProcessing (openems.cpp:535-680) instantiates the _CUDA class when the engine is an
 * Engine_CUDA (one `dynamic_cast<Engine_CUDA*>(FDTD_Eng)` per `new Process...`), see INTEGRATION.md.
 * Every class falls back to its base when the engine is not an Engine_CUDA or the dump type has no device path.
 */
#ifndef PROCESSING_CUDA_H
#define PROCESSING_CUDA_H

#include "Common/processcurrent.h"
#include "Common/processfields_td.h"
#include "Common/processfields_fd.h"
#include "Common/processmodematch.h"
#include "engine_cuda.h"

#include <vector>

class ProcessCurrent_CUDA : public ProcessCurrent
{
public:
	ProcessCurrent_CUDA(Engine_Interface_Base* eng_if) : ProcessCurrent(eng_if), m_probe(-1), m_slot(0) {}
	virtual double CalcIntegral();
protected:
	int m_probe;
	unsigned int m_slot;
	std::vector<double> m_values;
};

//! shared by the TD and FD dump classes: registers the dump box of a ProcessFields on the device
struct DeviceDumpBox
{
	DeviceDumpBox() : eng(NULL), dump_id(-1), ticket(-1), pinned(NULL) {}
	~DeviceDumpBox();
	bool Setup(Engine_Interface_Base* eng_if, const Operator_Base* op, int dump_type, const unsigned int numLines[3], unsigned int* const posLines[3]);
	Engine_CUDA* eng;
	int dump_id;
	long long ticket;
	float* pinned;       //!< page-locked {3,nz,ny,nx} block the asynchronous copy lands in
	size_t count;
};

class ProcessFieldsTD_CUDA : public ProcessFieldsTD
{
public:
	ProcessFieldsTD_CUDA(Engine_Interface_Base* eng_if) : ProcessFieldsTD(eng_if), m_pending(false), m_pending_ts(0), m_pending_time(0) {}
	virtual ~ProcessFieldsTD_CUDA();
	virtual void InitProcess();
	//! launches this timestep's dump asynchronously and writes out the PREVIOUS one: the time loop never waits
	virtual int Process();
	//! writes the last outstanding dump
	virtual void PostProcess();
protected:
	bool WriteOut(unsigned int ts, float time);
	DeviceDumpBox m_box;
	bool m_pending;
	unsigned int m_pending_ts;
	float m_pending_time;
};

class ProcessFieldsFD_CUDA : public ProcessFieldsFD
{
public:
	ProcessFieldsFD_CUDA(Engine_Interface_Base* eng_if) : ProcessFieldsFD(eng_if), m_fd_id(-1) {}
	virtual void InitProcess();
	virtual int Process();
	virtual void DumpFDData();
protected:
	DeviceDumpBox m_box;
	int m_fd_id;
};

class ProcessModeMatch_CUDA : public ProcessModeMatch
{
public:
	ProcessModeMatch_CUDA(Engine_Interface_Base* eng_if) : ProcessModeMatch(eng_if), m_eng(NULL), m_id(-1) {}
	virtual void InitProcess();
	virtual double* CalcMultipleIntegrals();
protected:
	Engine_CUDA* m_eng;
	int m_id;
};

#endif // PROCESSING_CUDA_H
