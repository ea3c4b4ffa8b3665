/*
 * engine_cuda.h -- Engine_CUDA: Engine (FDTD/engine.h) backed by libopenems_b200.so.
 * Goes to openEMS/FDTD/engine_cuda.h.  Compiled inside openEMS, see INTEGRATION.md.
 */
#ifndef ENGINE_CUDA_H
#define ENGINE_CUDA_H

#include "engine.h"
#include "openems_b200.h"

class Operator_CUDA;

class Engine_CUDA : public Engine
{
public:
	static Engine_CUDA* New(const Operator_CUDA* op);
	virtual ~Engine_CUDA();

	virtual void Init();
	virtual void Reset();

	//! only enqueues `iterTS` CUDA-graph launches; readers synchronise
	virtual bool IterateTS(unsigned int iterTS);
	virtual unsigned int GetNumberOfTimesteps();
	virtual void NextInterval(float curr_speed) {UNUSED(curr_speed);}

	// slow path for unknown callers (one 4-byte D2H/H2D per call)
	virtual FDTD_FLOAT GetVolt(unsigned int n, unsigned int x, unsigned int y, unsigned int z) const;
	virtual FDTD_FLOAT GetVolt(unsigned int n, const unsigned int pos[3]) const {return GetVolt(n,pos[0],pos[1],pos[2]);}
	virtual FDTD_FLOAT GetCurr(unsigned int n, unsigned int x, unsigned int y, unsigned int z) const;
	virtual FDTD_FLOAT GetCurr(unsigned int n, const unsigned int pos[3]) const {return GetCurr(n,pos[0],pos[1],pos[2]);}
	virtual void SetVolt(unsigned int n, unsigned int x, unsigned int y, unsigned int z, FDTD_FLOAT value);
	virtual void SetVolt(unsigned int n, const unsigned int pos[3], FDTD_FLOAT value) {SetVolt(n,pos[0],pos[1],pos[2],value);}
	virtual void SetCurr(unsigned int n, unsigned int x, unsigned int y, unsigned int z, FDTD_FLOAT value);
	virtual void SetCurr(unsigned int n, const unsigned int pos[3], FDTD_FLOAT value) {SetCurr(n,pos[0],pos[1],pos[2],value);}

	oems_cuda_engine* GetHandle() const {return m_h;}

protected:
	Engine_CUDA(const Operator_CUDA* op);
	//! maps every Operator_Extension to its device-side counterpart; refuses unknown ones
	virtual void InitExtensions();
	void Check(int rc, const char* what) const;

	const Operator_CUDA* m_Op_CUDA;
	class Engine_Ext_SteadyState* m_SSD; //!< stock extension object, kept only as the holder of GetLastDiff()
	oems_cuda_engine* m_h;
};

#endif // ENGINE_CUDA_H
