/* TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
 *
 * Bodies for the three reference classes whose own .cpp files need libraries that are not in this image
 * (boost::program_options, HDF5, VTK). The class DECLARATIONS are the reference's unmodified headers
 * (tools/global.h, tools/hdf5_file_writer.h, tools/vtk_file_writer.h); only the members are supplied here:
 *   - Global: verbose level 0, no command-line options (tools/global.cpp is the option parser, off the hot path);
 *   - HDF5_File_Writer / VTK_File_Writer: record what the reference's ProcessFields* classes write
 *     into an in-memory registry (ref_recorder) instead of a file, so that the dump path
 *     (Common/processfields_td.cpp:50-91, processfields_fd.cpp:110-230) can be observed by the tests.
 *     Dataset layout follows tools/hdf5_file_writer.cpp:286-302: {3, nz, ny, nx}, x fastest.
 */
#include "tools/global.h"
#include "tools/hdf5_file_writer.h"
#include "tools/vtk_file_writer.h"
#include "ref_recorder.h"

#include <map>
#include <mutex>

namespace po = boost::program_options;

// ---------------------------------------------------------------- Global (tools/global.h:29-137)
Global g_settings;

Global::Global()
{
	m_showProbeDiscretization = false;
	m_nativeFieldDumps = false;
	m_VerboseLevel = 0;
	m_SavedVerboseLevel = 0;
	m_optionDesc = NULL;
}
po::options_description Global::optionDesc() { return po::options_description(); }
void Global::appendOptionDesc(po::options_description) {}
void Global::clearOptionDesc() {}
void Global::parseLibraryArguments(std::vector<std::string>) {}
void Global::parseCommandLineArguments(int, const char**) {}
void Global::showOptionUsage(std::ostream&) {}
bool Global::hasOption(std::string) { return false; }
po::variable_value Global::getOption(std::string) { return po::variable_value(); }
void Global::clearOptions() {}

// ---------------------------------------------------------------- recorder
static std::mutex g_rec_mutex;
static std::map<std::string, ref_dataset> g_rec;

void ref_recorder_clear() { std::lock_guard<std::mutex> lk(g_rec_mutex); g_rec.clear(); }
void ref_recorder_put(const std::string& key, const ref_dataset& ds) { std::lock_guard<std::mutex> lk(g_rec_mutex); g_rec[key] = ds; }
const ref_dataset* ref_recorder_get(const std::string& key)
{
	std::lock_guard<std::mutex> lk(g_rec_mutex);
	std::map<std::string, ref_dataset>::const_iterator it = g_rec.find(key);
	return it == g_rec.end() ? NULL : &it->second;
}
std::vector<std::string> ref_recorder_keys()
{
	std::lock_guard<std::mutex> lk(g_rec_mutex);
	std::vector<std::string> out;
	for (std::map<std::string, ref_dataset>::const_iterator it = g_rec.begin(); it != g_rec.end(); ++it) out.push_back(it->first);
	return out;
}

// ---------------------------------------------------------------- HDF5_File_Writer (tools/hdf5_file_writer.h:27-69)
HDF5_File_Writer::HDF5_File_Writer(std::string filename) : m_filename(filename), m_Group("/") {}
HDF5_File_Writer::~HDF5_File_Writer() {}
hid_t HDF5_File_Writer::OpenGroup(hid_t, std::string) { return 0; }
void HDF5_File_Writer::SetCurrentGroup(std::string group, bool) { m_Group = group; }

bool HDF5_File_Writer::WriteRectMesh(unsigned int const* numLines, double const* const* discLines, int, double scaling)
{
	for (int n = 0; n < 3; ++n) {
		ref_dataset ds;
		ds.dims.push_back(numLines[n]);
		for (unsigned i = 0; i < numLines[n]; ++i) ds.data.push_back(discLines[n][i] * scaling);
		ref_recorder_put(m_filename + ":/Mesh/" + std::string(1, (char)('x' + n)), ds);
	}
	return true;
}
bool HDF5_File_Writer::WriteRectMesh(unsigned int const* numLines, float const* const* discLines, int, float scaling)
{
	for (int n = 0; n < 3; ++n) {
		ref_dataset ds;
		ds.dims.push_back(numLines[n]);
		for (unsigned i = 0; i < numLines[n]; ++i) ds.data.push_back(discLines[n][i] * scaling);
		ref_recorder_put(m_filename + ":/Mesh/" + std::string(1, (char)('x' + n)), ds);
	}
	return true;
}

template <typename T>
static bool rec_scalar(const std::string& key, T const* const* const* field, size_t datasize[3])
{
	ref_dataset ds;
	ds.dims = { datasize[2], datasize[1], datasize[0] };
	ds.data.resize(datasize[0] * datasize[1] * datasize[2]);
	size_t pos = 0;
	for (size_t k = 0; k < datasize[2]; ++k)
		for (size_t j = 0; j < datasize[1]; ++j)
			for (size_t i = 0; i < datasize[0]; ++i) ds.data[pos++] = field[i][j][k];
	ref_recorder_put(key, ds);
	return true;
}
template <typename T>
static bool rec_vector(const std::string& key, T const* const* const* const* field, size_t datasize[3])
{
	ref_dataset ds;
	ds.dims = { 3, datasize[2], datasize[1], datasize[0] };
	ds.data.resize(3 * datasize[0] * datasize[1] * datasize[2]);
	size_t pos = 0;
	for (int n = 0; n < 3; ++n)
		for (size_t k = 0; k < datasize[2]; ++k)
			for (size_t j = 0; j < datasize[1]; ++j)
				for (size_t i = 0; i < datasize[0]; ++i) ds.data[pos++] = field[n][i][j][k];
	ref_recorder_put(key, ds);
	return true;
}
template <typename T>
static bool rec_cvector(const std::string& key, std::complex<T> const* const* const* const* field, size_t datasize[3])
{
	ref_dataset re, im;
	re.dims = im.dims = { 3, datasize[2], datasize[1], datasize[0] };
	re.data.resize(3 * datasize[0] * datasize[1] * datasize[2]);
	im.data.resize(re.data.size());
	size_t pos = 0;
	for (int n = 0; n < 3; ++n)
		for (size_t k = 0; k < datasize[2]; ++k)
			for (size_t j = 0; j < datasize[1]; ++j)
				for (size_t i = 0; i < datasize[0]; ++i, ++pos) { re.data[pos] = field[n][i][j][k].real(); im.data[pos] = field[n][i][j][k].imag(); }
	ref_recorder_put(key + "_real", re);
	ref_recorder_put(key + "_imag", im);
	return true;
}

bool HDF5_File_Writer::WriteScalarField(std::string n, float const* const* const* f, size_t d[3]) { return rec_scalar(m_filename + ":" + m_Group + "/" + n, f, d); }
bool HDF5_File_Writer::WriteScalarField(std::string n, double const* const* const* f, size_t d[3]) { return rec_scalar(m_filename + ":" + m_Group + "/" + n, f, d); }
bool HDF5_File_Writer::WriteScalarField(std::string, std::complex<float> const* const* const*, size_t[3]) { return true; }
bool HDF5_File_Writer::WriteScalarField(std::string, std::complex<double> const* const* const*, size_t[3]) { return true; }
bool HDF5_File_Writer::WriteVectorField(std::string n, float const* const* const* const* f, size_t d[3]) { return rec_vector(m_filename + ":" + m_Group + "/" + n, f, d); }
bool HDF5_File_Writer::WriteVectorField(std::string n, double const* const* const* const* f, size_t d[3]) { return rec_vector(m_filename + ":" + m_Group + "/" + n, f, d); }
bool HDF5_File_Writer::WriteVectorField(std::string n, std::complex<float> const* const* const* const* f, size_t d[3]) { return rec_cvector(m_filename + ":" + m_Group + "/" + n, f, d); }
bool HDF5_File_Writer::WriteVectorField(std::string n, std::complex<double> const* const* const* const* f, size_t d[3]) { return rec_cvector(m_filename + ":" + m_Group + "/" + n, f, d); }

template <typename T>
static bool rec_flat(const std::string& key, T const* buf, size_t dim, size_t* datasize)
{
	ref_dataset ds;
	size_t n = 1;
	for (size_t d = 0; d < dim; ++d) { ds.dims.push_back(datasize[d]); n *= datasize[d]; }
	ds.data.assign(buf, buf + n);
	ref_recorder_put(key, ds);
	return true;
}
bool HDF5_File_Writer::WriteData(std::string n, float const* b, size_t dim, size_t* d) { return rec_flat(m_filename + ":" + m_Group + "/" + n, b, dim, d); }
bool HDF5_File_Writer::WriteData(std::string n, double const* b, size_t dim, size_t* d) { return rec_flat(m_filename + ":" + m_Group + "/" + n, b, dim, d); }
bool HDF5_File_Writer::WriteData(std::string, hid_t, void const*, size_t, size_t*) { return true; }

bool HDF5_File_Writer::WriteAtrribute(std::string, std::string, void const*, hsize_t, hid_t) { return true; }
bool HDF5_File_Writer::WriteAtrribute(std::string loc, std::string name, float const* v, hsize_t size)
{
	size_t sz = (size_t)size;
	return rec_flat(m_filename + ":" + loc + "@" + name, v, 1, &sz);
}
bool HDF5_File_Writer::WriteAtrribute(std::string loc, std::string name, double const* v, hsize_t size)
{
	size_t sz = (size_t)size;
	return rec_flat(m_filename + ":" + loc + "@" + name, v, 1, &sz);
}
bool HDF5_File_Writer::WriteAtrribute(std::string loc, std::string name, std::vector<float> v) { return WriteAtrribute(loc, name, v.data(), v.size()); }
bool HDF5_File_Writer::WriteAtrribute(std::string loc, std::string name, std::vector<double> v) { return WriteAtrribute(loc, name, v.data(), v.size()); }
bool HDF5_File_Writer::WriteAtrribute(std::string loc, std::string name, float v) { return WriteAtrribute(loc, name, &v, 1); }
bool HDF5_File_Writer::WriteAtrribute(std::string loc, std::string name, double v) { return WriteAtrribute(loc, name, &v, 1); }

// ---------------------------------------------------------------- VTK_File_Writer (tools/vtk_file_writer.h:33-104)
// fields are recorded on Write() under "<filename>[_<timestep>]:<fieldname>", {3, nz, ny, nx}
struct vtk_pending { std::string name; ref_dataset ds; };
static std::map<const VTK_File_Writer*, std::vector<vtk_pending> > g_vtk_pending;

VTK_File_Writer::VTK_File_Writer(std::string filename, int meshType)
{
	SetFilename(filename);
	m_MeshType = meshType;
	m_NativeDump = false;
	m_Binary = true;
	m_Compress = true;
	m_AppendMode = false;
	m_ActiveTS = false;
	m_timestep = 0;
	m_GridData = NULL;
}
VTK_File_Writer::~VTK_File_Writer() { g_vtk_pending.erase(this); }
void VTK_File_Writer::SetMeshLines(double const* const* lines, unsigned int const* count, double scaling)
{
	for (int n = 0; n < 3; ++n) {
		m_MeshLines[n].clear();
		for (unsigned i = 0; i < count[n]; ++i) m_MeshLines[n].push_back(lines[n][i] * scaling);
	}
}
template <typename T>
static void vtk_add_vec(const VTK_File_Writer* w, const std::vector<double>* ml, const std::string& name, T const* const* const* const* f)
{
	size_t d[3] = { ml[0].size(), ml[1].size(), ml[2].size() };
	vtk_pending p;
	p.name = name;
	p.ds.dims = { 3, d[2], d[1], d[0] };
	p.ds.data.resize(3 * d[0] * d[1] * d[2]);
	size_t pos = 0;
	for (int n = 0; n < 3; ++n)
		for (size_t k = 0; k < d[2]; ++k)
			for (size_t j = 0; j < d[1]; ++j)
				for (size_t i = 0; i < d[0]; ++i) p.ds.data[pos++] = f[n][i][j][k];
	g_vtk_pending[w].push_back(p);
}
void VTK_File_Writer::AddScalarField(std::string, double const* const* const*) {}
void VTK_File_Writer::AddScalarField(std::string, float const* const* const*) {}
void VTK_File_Writer::AddVectorField(std::string name, double const* const* const* const* f) { vtk_add_vec(this, m_MeshLines, name, f); }
void VTK_File_Writer::AddVectorField(std::string name, float const* const* const* const* f) { vtk_add_vec(this, m_MeshLines, name, f); }
void VTK_File_Writer::AddVectorField(std::string name, ArrayLib::ArrayNIJK<float>& f)
{
	size_t d[3] = { m_MeshLines[0].size(), m_MeshLines[1].size(), m_MeshLines[2].size() };
	vtk_pending p;
	p.name = name;
	p.ds.dims = { 3, d[2], d[1], d[0] };
	p.ds.data.resize(3 * d[0] * d[1] * d[2]);
	size_t pos = 0;
	for (unsigned n = 0; n < 3; ++n)
		for (unsigned k = 0; k < d[2]; ++k)
			for (unsigned j = 0; j < d[1]; ++j)
				for (unsigned i = 0; i < d[0]; ++i) p.ds.data[pos++] = f(n, i, j, k);
	g_vtk_pending[this].push_back(p);
}
int VTK_File_Writer::GetNumberOfFields() const
{
	std::map<const VTK_File_Writer*, std::vector<vtk_pending> >::const_iterator it = g_vtk_pending.find(this);
	return it == g_vtk_pending.end() ? 0 : (int)it->second.size();
}
void VTK_File_Writer::ClearAllFields() { g_vtk_pending[this].clear(); }
std::string VTK_File_Writer::GetTimestepFilename(int pad_length) const
{
	if (!m_ActiveTS) return m_filename;
	std::string ts = std::to_string(m_timestep);
	while ((int)ts.size() < pad_length) ts = "0" + ts;
	return m_filename + "_" + ts;
}
bool VTK_File_Writer::Write()
{
	std::vector<vtk_pending>& v = g_vtk_pending[this];
	for (size_t i = 0; i < v.size(); ++i) ref_recorder_put(GetTimestepFilename() + ":" + v[i].name, v[i].ds);
	return true;
}
bool VTK_File_Writer::WriteASCII() { return Write(); }
bool VTK_File_Writer::WriteXML() { return Write(); }
