/*
 * operator_cuda.h -- Operator_CUDA: the operator half of --engine=cuda.
 * Goes to openEMS/FDTD/operator_cuda.h.  Compiled inside openEMS; in this repository it is compiled
 * against the reference's unmodified headers and run by the oracle/_ref harness
 * (oracle/Makefile.ref, target libopenems_ref_cuda.so; tests/test_gpu_reference_integration.py).
 *
 * Derives from Operator_Multithread so the host build (Calc_EC, CalcPEC, extensions'
 * BuildExtension) stays the reference's own, threaded code; only CreateEngine differs.
 * (The sse compression of Operator_SSE_Compressed::CalcECOperator still runs -- it cannot be
 * skipped without editing the reference -- and is simply not used: Engine_CUDA reads the
 * coefficients through GetVV/GetVI/GetII/GetIV and the library re-keys them per cell.)
 */
#ifndef OPERATOR_CUDA_H
#define OPERATOR_CUDA_H

#include "operator_multithread.h"

class Operator_CUDA : public Operator_Multithread
{
	friend class Engine_CUDA;
public:
	static Operator_CUDA* New(unsigned int numThreads = 0, int device = -1);
	virtual ~Operator_CUDA() {}

	//! returns an Engine_CUDA; the single host->device crossing happens in there
	virtual Engine* CreateEngine();

	int GetDevice() const {return m_device;}

protected:
	Operator_CUDA() : Operator_Multithread(), m_device(-1) {}
	int m_device;
};

#endif // OPERATOR_CUDA_H
