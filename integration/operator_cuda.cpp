/* operator_cuda.cpp -- see operator_cuda.h.  Goes to openEMS/FDTD/operator_cuda.cpp. */
#include "operator_cuda.h"
#include "engine_cuda.h"

using std::cout;
using std::endl;

Operator_CUDA* Operator_CUDA::New(unsigned int numThreads, int device)
{
	cout << "Create FDTD operator (B200 CUDA engine, host build multi-threaded)" << endl;
	Operator_CUDA* op = new Operator_CUDA();
	op->setNumThreads(numThreads);
	op->m_device = device;
	op->Init();
	return op;
}

Engine* Operator_CUDA::CreateEngine()
{
	m_Engine = Engine_CUDA::New(this);
	return m_Engine;
}
