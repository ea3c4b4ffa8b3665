/*
 * operator_cuda.h -- Operator_CUDA: the operator half of --engine=cuda.
 * Goes to openEMS/FDTD/operator_cuda.h.  Compiled inside openEMS (needs its headers and
 * CSXCAD); it is NOT built in this repository, see INTEGRATION.md.
 *
 * Derives from Operator_Multithread so the host build (Calc_EC, CalcPEC, extensions'
 * BuildExtension) stays the reference's own, threaded code; only CreateEngine differs.
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
	//! keep the dense f4 arrays: Engine_CUDA re-keys them per cell (library side), so the
	//! SSE compression (per 4 interleaved z cells, operator_sse_compressed.cpp:114-175) is skipped
	virtual int CalcECOperator( DebugFlags debugFlags = None );
	int m_device;
};

#endif // OPERATOR_CUDA_H
