/* TEST INFRASTRUCTURE shim: no-op stand-ins for the four VTK classes FDTD/operator.cpp:644-733 uses in its
 * debug dump (Operator::DumpPEC2File). The dump is never requested by the oracle/_ref driver. */
#pragma once
#define VTK_MAJOR_VERSION 9
struct vtkObjectStub { void Delete() { delete this; } virtual ~vtkObjectStub() {} };
struct vtkPoints : vtkObjectStub { static vtkPoints* New() { return new vtkPoints; } long long InsertNextPoint(const double*) { return m_n++; } long long m_n = 0; };
struct vtkCellArray : vtkObjectStub { static vtkCellArray* New() { return new vtkCellArray; } void InsertNextCell(int) {} void InsertCellPoint(long long) {} };
struct vtkPolyData : vtkObjectStub { static vtkPolyData* New() { return new vtkPolyData; } void SetPoints(vtkPoints*) {} void SetLines(vtkCellArray*) {} };
struct vtkXMLPolyDataWriter : vtkObjectStub { static vtkXMLPolyDataWriter* New() { return new vtkXMLPolyDataWriter; }
	void SetFileName(const char*) {} void SetInputData(vtkPolyData*) {} void SetInput(vtkPolyData*) {} int Write() { return 0; } };
