/* operator_cuda.cpp -- see operator_cuda.h.  Goes to openEMS/FDTD/operator_cuda.cpp. */
#include "operator_cuda.h"
#include "engine_cuda.h"

Operator_CUDA* Operator_CUDA::New(unsigned int numThreads, int device)
{
	cout << "Create FDTD operator (B200 CUDA engine, host build multi-threaded)" << endl;
	Operator_CUDA* op = new Operator_CUDA();
	op->setNumThreads(numThreads);
	op->m_device = device;
	op->Init();
	return op;
}

int Operator_CUDA::CalcECOperator( DebugFlags debugFlags )
{
	// Operator_SSE_Compressed::CalcECOperator would compress after the build
	// (operator_sse_compressed.cpp:56-63); switch that off and run the threaded build.
	m_Use_Compression = false;
	m_max_fifo = 0;
	return Operator_Multithread::CalcECOperator( debugFlags );
}

Engine* Operator_CUDA::CreateEngine()
{
	m_Engine = Engine_CUDA::New(this);
	return m_Engine;
}
