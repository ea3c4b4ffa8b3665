/* TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
 *
 * A minimal stand-in for the CSXCAD geometry library (thliebig/CSXCAD, branch master, unpinned:
 * .github/scripts/resolve_dependent_repos.py:74-78), which is NOT vendored under /root/reference.
 * It exists so that the reference's own, UNMODIFIED translation units (FDTD/operator.cpp, the
 * FDTD/extensions/operator_ext_*.cpp builders, Common/processing.cpp ...) compile and run here.
 * Only what those TUs call is provided, written from their call sites (file:line given per item):
 *   - a rectilinear grid, axis-aligned boxes and curves as the only primitives,
 *   - constant (unweighted) material / excitation parameters,
 *   - priority lookup: highest priority wins, for equal priority the primitive added LAST wins
 *     (CSXCAD sorts its primitive list by priority, later IDs first on ties).
 * Geometry-driven coefficients therefore stay "parity unpinned" for anything beyond boxes
 * (SURVEY.md section 8c); the time loop, the operator arithmetic and the extension builders that run on top
 * of it are the reference's own code.
 */
#ifndef CSXCAD_MINI_H
#define CSXCAD_MINI_H

#include <string>
#include <vector>
#include <iostream>
#include <algorithm>
#include <cmath>
#include <cstring>

#ifndef CSXCAD_EXPORT
#define CSXCAD_EXPORT
#endif

enum CoordinateSystem { CARTESIAN = 0, CYLINDRICAL = 1, UNDEFINED_CS = 2 };

//! Cartesian <-> cylindrical transform (CSXCAD CSUseful); FDTD/operator.cpp:208,689
inline double* TransformCoordSystem(const double* in, double* out, CoordinateSystem cs_in, CoordinateSystem cs_out)
{
	double t[3] = { in[0], in[1], in[2] };
	if (cs_in == CYLINDRICAL && (cs_out == CARTESIAN)) { t[0] = in[0] * cos(in[1]); t[1] = in[0] * sin(in[1]); }
	else if (cs_in == CARTESIAN && cs_out == CYLINDRICAL) { t[0] = sqrt(in[0]*in[0] + in[1]*in[1]); t[1] = atan2(in[1], in[0]); }
	out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
	return out;
}

//! foot point parameter (0 at start, 1 at stop) and distance of P to the line start->stop; FDTD/operator.cpp:409-470
inline void Point_Line_Distance(const double P[], const double start[], const double stop[], double& foot, double& dist, CoordinateSystem = UNDEFINED_CS)
{
	double dir[3] = { stop[0]-start[0], stop[1]-start[1], stop[2]-start[2] };
	double LL = dir[0]*dir[0] + dir[1]*dir[1] + dir[2]*dir[2];
	if (LL == 0) { foot = 0; dist = sqrt((P[0]-start[0])*(P[0]-start[0]) + (P[1]-start[1])*(P[1]-start[1]) + (P[2]-start[2])*(P[2]-start[2])); return; }
	foot = ((P[0]-start[0])*dir[0] + (P[1]-start[1])*dir[1] + (P[2]-start[2])*dir[2]) / LL;
	double d[3] = { P[0]-(start[0]+foot*dir[0]), P[1]-(start[1]+foot*dir[1]), P[2]-(start[2]+foot*dir[2]) };
	dist = sqrt(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]);
}

//! CSXCAD CSUseful.h: integer to string (FDTD/operator_cylindermultigrid.cpp:326)
inline std::string ConvertInt(int number) { return std::to_string(number); }

class ParameterSet
{
public:
	ParameterSet() {}
};

//! a coordinate triple; GetCoords(cs) as used by FDTD/operator.cpp:1634, operator_ext_lumpedRLC.cpp:263
class ParameterCoord
{
public:
	ParameterCoord() { m_c[0] = m_c[1] = m_c[2] = 0; }
	void SetValue(int n, double v) { m_c[n] = v; }
	double GetValue(int n) const { return m_c[n]; }
	const double* GetNativeCoords() const { return m_c; }
	const double* GetCartesianCoords() const { return m_c; }
	const double* GetCoords(CoordinateSystem) const { return m_c; }
protected:
	double m_c[3];
};

//! rectilinear grid: FDTD/operator.cpp:793-836 (GetLines, GetDeltaUnit, Clone), openems.cpp
class CSRectGrid
{
public:
	CSRectGrid() : m_deltaUnit(1), m_meshType(CARTESIAN) {}
	static CSRectGrid* Clone(CSRectGrid* g) { return new CSRectGrid(*g); }
	void AddDiscLine(int n, double v) { m_lines[n].push_back(v); }
	void AddDiscLines(int n, int cnt, const double* v) { for (int i = 0; i < cnt; ++i) m_lines[n].push_back(v[i]); }
	void ClearLines(int n) { m_lines[n].clear(); }
	void SetDeltaUnit(double d) { m_deltaUnit = d; }
	double GetDeltaUnit() const { return m_deltaUnit; }
	void Sort(int n)
	{
		std::sort(m_lines[n].begin(), m_lines[n].end());
		m_lines[n].erase(std::unique(m_lines[n].begin(), m_lines[n].end()), m_lines[n].end());
	}
	size_t GetQtyLines(int n) const { return m_lines[n].size(); }
	double GetLine(int n, size_t i) const { return m_lines[n].at(i); }
	bool SetLine(int n, size_t i, double v) { if (i >= m_lines[n].size()) return false; m_lines[n][i] = v; return true; }
	//! returns a new[] array (the caller's previous array is deleted), CSXCAD semantics
	double* GetLines(int n, double* array, unsigned int& qty, bool sorted = true)
	{
		if (sorted) Sort(n);
		delete[] array;
		qty = (unsigned int)m_lines[n].size();
		array = new double[qty];
		for (unsigned int i = 0; i < qty; ++i) array[i] = m_lines[n][i];
		return array;
	}
	void SetMeshType(CoordinateSystem t) { m_meshType = t; }
	CoordinateSystem GetMeshType() const { return m_meshType; }
	double* GetSimArea()
	{
		for (int n = 0; n < 3; ++n) {
			Sort(n);
			m_simBox[2*n] = m_lines[n].empty() ? 0 : m_lines[n].front();
			m_simBox[2*n+1] = m_lines[n].empty() ? 0 : m_lines[n].back();
		}
		return m_simBox;
	}
protected:
	std::vector<double> m_lines[3];
	double m_deltaUnit;
	CoordinateSystem m_meshType;
	double m_simBox[6];
};

class CSProperties;
class CSPrimBox;
class CSPrimCurve;
class CSPropMaterial;
class CSPropExcitation;
class CSPropLorentzMaterial;
class CSPropDebyeMaterial;

class CSPrimitives
{
public:
	enum PrimitiveType { POINT, BOX, MULTIBOX, SPHERE, SPHERICALSHELL, CYLINDER, CYLINDRICALSHELL, POLYGON,
	                     LINPOLY, ROTPOLY, POLYHEDRON, CURVE, WIRE, USERDEFINED, POLYHEDRONREADER };
	virtual ~CSPrimitives() {}
	unsigned int GetID() const { return m_id; }
	void SetID(unsigned int id) { m_id = id; }
	PrimitiveType GetType() const { return m_type; }
	std::string GetTypeName() const { return m_type == BOX ? "Box" : (m_type == CURVE ? "Curve" : "Primitive"); }
	int GetPriority() const { return m_priority; }
	void SetPriority(int p) { m_priority = p; }
	CSProperties* GetProperty() const { return m_prop; }
	void SetPrimitiveUsed(bool v) { m_used = v; }
	bool GetPrimitiveUsed() const { return m_used; }
	CoordinateSystem GetCoordinateSystem() const { return CARTESIAN; }
	CoordinateSystem GetBoundBoxCoordSystem() const { return CARTESIAN; }
	virtual bool GetBoundBox(double box[6], bool = false) { for (int i = 0; i < 6; ++i) box[i] = m_bb[i]; return true; }
	//! number of non-degenerate axes of the bounding box (operator_ext_conductingsheet.cpp:102)
	virtual int GetDimension()
	{
		int d = 0;
		for (int n = 0; n < 3; ++n) if (m_bb[2*n] != m_bb[2*n+1]) ++d;
		return d;
	}
	virtual bool IsInside(const double* coord, double tol = 0) = 0;
	virtual CSPrimBox* ToBox() { return NULL; }
	virtual CSPrimCurve* ToCurve() { return NULL; }
protected:
	CSPrimitives(PrimitiveType t, CSProperties* prop);
	unsigned int m_id;
	PrimitiveType m_type;
	int m_priority;
	CSProperties* m_prop;
	bool m_used;
	double m_bb[6];
};

//! axis-aligned box, inclusive containment on all faces
class CSPrimBox : public CSPrimitives
{
public:
	CSPrimBox(CSProperties* prop, const double start[3], const double stop[3], int priority = 0) : CSPrimitives(BOX, prop)
	{
		m_priority = priority;
		for (int n = 0; n < 3; ++n) {
			m_start.SetValue(n, start[n]);
			m_stop.SetValue(n, stop[n]);
			m_bb[2*n] = std::min(start[n], stop[n]);
			m_bb[2*n+1] = std::max(start[n], stop[n]);
		}
	}
	ParameterCoord* GetStartCoord() { return &m_start; }
	ParameterCoord* GetStopCoord() { return &m_stop; }
	virtual bool IsInside(const double* c, double tol = 0)
	{
		for (int n = 0; n < 3; ++n)
			if ((m_bb[2*n] - tol > c[n]) || (m_bb[2*n+1] + tol < c[n])) return false;
		return true;
	}
	virtual CSPrimBox* ToBox() { return this; }
protected:
	ParameterCoord m_start, m_stop;
};

//! polyline of points; never "inside" anything (a curve has no volume), used through Operator::FindPath
class CSPrimCurve : public CSPrimitives
{
public:
	CSPrimCurve(CSProperties* prop, int priority = 0) : CSPrimitives(CURVE, prop)
	{
		m_priority = priority;
		for (int i = 0; i < 6; ++i) m_bb[i] = 0;
	}
	size_t AddPoint(const double p[3])
	{
		for (int n = 0; n < 3; ++n) {
			m_pts[n].push_back(p[n]);
			if (m_pts[n].size() == 1) m_bb[2*n] = m_bb[2*n+1] = p[n];
			m_bb[2*n] = std::min(m_bb[2*n], p[n]);
			m_bb[2*n+1] = std::max(m_bb[2*n+1], p[n]);
		}
		return m_pts[0].size();
	}
	size_t GetNumberOfPoints() const { return m_pts[0].size(); }
	bool GetPoint(size_t i, double p[3], bool = false) const
	{
		if (i >= m_pts[0].size()) return false;
		for (int n = 0; n < 3; ++n) p[n] = m_pts[n][i];
		return true;
	}
	bool GetPoint(size_t i, double p[3], CoordinateSystem, bool = false) const { return GetPoint(i, p); }
	virtual bool IsInside(const double*, double = 0) { return false; }
	virtual CSPrimCurve* ToCurve() { return this; }
protected:
	std::vector<double> m_pts[3];
};

class CSProperties
{
public:
	//! bit masks as in CSXCAD (types are OR-combined along the class hierarchy)
	enum PropertyType
	{
		ANY = 0xfff, UNKNOWN = 0x001, MATERIAL = 0x002, METAL = 0x004, EXCITATION = 0x008, PROBEBOX = 0x010,
		RESBOX = 0x020, DUMPBOX = 0x040, /* unused = 0x080, */ DISPERSIVEMATERIAL = 0x100, LORENTZMATERIAL = 0x200,
		DEBYEMATERIAL = 0x400, DISCRETE_MATERIAL = 0x1000, LUMPED_ELEMENT = 0x2000, CONDUCTINGSHEET = 0x4000,
		ABSORBING_BC = 0x8000
	};
	virtual ~CSProperties() { for (size_t i = 0; i < m_prims.size(); ++i) delete m_prims[i]; }
	int GetType() const { return m_type; }
	const std::string GetName() const { return m_name; }
	void SetName(const std::string& n) { m_name = n; }
	unsigned int GetID() const { return m_id; }
	void SetID(unsigned int id) { m_id = id; }
	bool GetMaterial() const { return m_isMaterial; }
	size_t GetQtyPrimitives() const { return m_prims.size(); }
	CSPrimitives* GetPrimitive(size_t n) { return n < m_prims.size() ? m_prims[n] : NULL; }
	std::vector<CSPrimitives*> GetAllPrimitives(bool = false) { return m_prims; }
	void AddPrimitive(CSPrimitives* p) { m_prims.push_back(p); }
	virtual CSPropMaterial* ToMaterial() { return NULL; }
	virtual CSPropExcitation* ToExcitation() { return NULL; }
	virtual CSPropLorentzMaterial* ToLorentzMaterial() { return NULL; }
	virtual CSPropDebyeMaterial* ToDebyeMaterial() { return NULL; }
	const std::string GetTypeString() const { return "property"; }
	void WarnUnusedPrimitves(std::ostream&) {}
protected:
	CSProperties(int type) : m_type(type), m_id(0), m_isMaterial(false) {}
	int m_type;
	unsigned int m_id;
	bool m_isMaterial;
	std::string m_name;
	std::vector<CSPrimitives*> m_prims;
};

inline CSPrimitives::CSPrimitives(PrimitiveType t, CSProperties* prop) : m_id(0), m_type(t), m_priority(0), m_prop(prop), m_used(false)
{
	if (prop) prop->AddPrimitive(this);
}

class CSPropMetal : public CSProperties
{
public:
	CSPropMetal() : CSProperties(METAL) { m_isMaterial = true; }
protected:
	CSPropMetal(int type) : CSProperties(type | METAL) { m_isMaterial = true; }
};

//! FDTD/extensions/operator_ext_conductingsheet.cpp:96-132
class CSPropConductingSheet : public CSPropMetal
{
public:
	CSPropConductingSheet(double conductivity, double thickness) : CSPropMetal(CONDUCTINGSHEET), m_cond(conductivity), m_thick(thickness) {}
	double GetConductivity() const { return m_cond; }
	double GetThickness() const { return m_thick; }
protected:
	double m_cond, m_thick;
};

//! constant, optionally anisotropic material. *Weighted getters: FDTD/operator.cpp:1297-1311
class CSPropMaterial : public CSProperties
{
public:
	CSPropMaterial() : CSProperties(MATERIAL) { init(); }
	void SetEpsilon(double v, int ny = -1) { set(m_eps, v, ny); }
	void SetMue(double v, int ny = -1) { set(m_mue, v, ny); }
	void SetKappa(double v, int ny = -1) { set(m_kappa, v, ny); }
	void SetSigma(double v, int ny = -1) { set(m_sigma, v, ny); }
	void SetDensity(double v) { m_density = v; }
	double GetEpsilon(int ny = 0) const { return m_eps[ny]; }
	double GetMue(int ny = 0) const { return m_mue[ny]; }
	double GetKappa(int ny = 0) const { return m_kappa[ny]; }
	double GetSigma(int ny = 0) const { return m_sigma[ny]; }
	double GetDensity() const { return m_density; }
	double GetEpsilonWeighted(int ny, const double*) { return m_eps[ny]; }
	double GetMueWeighted(int ny, const double*) { return m_mue[ny]; }
	double GetKappaWeighted(int ny, const double*) { return m_kappa[ny]; }
	double GetSigmaWeighted(int ny, const double*) { return m_sigma[ny]; }
	double GetDensityWeighted(const double*) { return m_density; }
	bool GetIsotropy() const { return m_eps[0] == m_eps[1] && m_eps[1] == m_eps[2] && m_kappa[0] == m_kappa[1] && m_kappa[1] == m_kappa[2]; }
	virtual CSPropMaterial* ToMaterial() { return this; }
protected:
	CSPropMaterial(int type) : CSProperties(type | MATERIAL) { init(); }
	void init()
	{
		m_isMaterial = true;
		for (int n = 0; n < 3; ++n) { m_eps[n] = 1; m_mue[n] = 1; m_kappa[n] = 0; m_sigma[n] = 0; }
		m_density = 0;
	}
	static void set(double* a, double v, int ny) { if (ny < 0) a[0] = a[1] = a[2] = v; else a[ny] = v; }
	double m_eps[3], m_mue[3], m_kappa[3], m_sigma[3], m_density;
};

class CSPropDispersiveMaterial : public CSPropMaterial
{
public:
	int GetDispersionOrder() const { return m_order; }
	void SetDispersionOrder(int order) { m_order = order; resize(); }
protected:
	CSPropDispersiveMaterial(int type) : CSPropMaterial(type | DISPERSIVEMATERIAL), m_order(0) {}
	virtual void resize() {}
	int m_order;
};

//! Drude/Lorentz poles: FDTD/extensions/operator_ext_lorentzmaterial.cpp:240-307
class CSPropLorentzMaterial : public CSPropDispersiveMaterial
{
public:
	CSPropLorentzMaterial() : CSPropDispersiveMaterial(LORENTZMATERIAL) { SetDispersionOrder(1); }
	void SetEpsPlasmaFreq(int o, double v, int ny = -1) { setv(m_epsPlasma, o, v, ny); }
	void SetEpsLorPoleFreq(int o, double v, int ny = -1) { setv(m_epsPole, o, v, ny); }
	void SetEpsRelaxTime(int o, double v, int ny = -1) { setv(m_epsRelax, o, v, ny); }
	void SetMuePlasmaFreq(int o, double v, int ny = -1) { setv(m_muePlasma, o, v, ny); }
	void SetMueLorPoleFreq(int o, double v, int ny = -1) { setv(m_muePole, o, v, ny); }
	void SetMueRelaxTime(int o, double v, int ny = -1) { setv(m_mueRelax, o, v, ny); }
	double GetEpsPlasmaFreqWeighted(int o, int ny, const double*) { return getv(m_epsPlasma, o, ny); }
	double GetEpsLorPoleFreqWeighted(int o, int ny, const double*) { return getv(m_epsPole, o, ny); }
	double GetEpsRelaxTimeWeighted(int o, int ny, const double*) { return getv(m_epsRelax, o, ny); }
	double GetMuePlasmaFreqWeighted(int o, int ny, const double*) { return getv(m_muePlasma, o, ny); }
	double GetMueLorPoleFreqWeighted(int o, int ny, const double*) { return getv(m_muePole, o, ny); }
	double GetMueRelaxTimeWeighted(int o, int ny, const double*) { return getv(m_mueRelax, o, ny); }
	virtual CSPropLorentzMaterial* ToLorentzMaterial() { return this; }
protected:
	typedef std::vector<double> vd;
	virtual void resize()
	{
		vd* all[6] = { m_epsPlasma, m_epsPole, m_epsRelax, m_muePlasma, m_muePole, m_mueRelax };
		for (int a = 0; a < 6; ++a) for (int n = 0; n < 3; ++n) all[a][n].resize(m_order, 0.0);
	}
	static void setv(vd* a, int o, double v, int ny) { if (ny < 0) { for (int n = 0; n < 3; ++n) a[n].at(o) = v; } else a[ny].at(o) = v; }
	static double getv(vd* a, int o, int ny) { return (o >= 0 && o < (int)a[ny].size()) ? a[ny][o] : 0; }
	vd m_epsPlasma[3], m_epsPole[3], m_epsRelax[3], m_muePlasma[3], m_muePole[3], m_mueRelax[3];
};

//! Debye poles: FDTD/extensions/operator_ext_lorentzmaterial.cpp:259-262
class CSPropDebyeMaterial : public CSPropDispersiveMaterial
{
public:
	CSPropDebyeMaterial() : CSPropDispersiveMaterial(DEBYEMATERIAL) { SetDispersionOrder(1); }
	void SetEpsDelta(int o, double v, int ny = -1) { setv(m_epsDelta, o, v, ny); }
	void SetEpsRelaxTime(int o, double v, int ny = -1) { setv(m_epsRelax, o, v, ny); }
	double GetEpsDeltaWeighted(int o, int ny, const double*) { return getv(m_epsDelta, o, ny); }
	double GetEpsRelaxTimeWeighted(int o, int ny, const double*) { return getv(m_epsRelax, o, ny); }
	virtual CSPropDebyeMaterial* ToDebyeMaterial() { return this; }
protected:
	typedef std::vector<double> vd;
	virtual void resize() { for (int n = 0; n < 3; ++n) { m_epsDelta[n].resize(m_order, 0.0); m_epsRelax[n].resize(m_order, 0.0); } }
	static void setv(vd* a, int o, double v, int ny) { if (ny < 0) { for (int n = 0; n < 3; ++n) a[n].at(o) = v; } else a[ny].at(o) = v; }
	static double getv(vd* a, int o, int ny) { return (o >= 0 && o < (int)a[ny].size()) ? a[ny][o] : 0; }
	vd m_epsDelta[3], m_epsRelax[3];
};

//! FDTD/operator.cpp:1586-1763, FDTD/extensions/operator_ext_lumpedRLC.cpp:164-444
class CSPropLumpedElement : public CSProperties
{
public:
	enum LEtype { PARALLEL = 0, SERIES = 1, INVALID = -1 };
	CSPropLumpedElement() : CSProperties(LUMPED_ELEMENT), m_R(NAN), m_C(NAN), m_L(NAN), m_ny(-1), m_caps(true), m_LEtype(PARALLEL) {}
	void SetResistance(double v) { m_R = v; }
	void SetCapacity(double v) { m_C = v; }
	void SetInductance(double v) { m_L = v; }
	void SetDirection(int ny) { m_ny = ny; }
	void SetCaps(bool v) { m_caps = v; }
	void SetLEtype(LEtype t) { m_LEtype = t; }
	double GetResistance() const { return m_R; }
	double GetCapacity() const { return m_C; }
	double GetInductance() const { return m_L; }
	int GetDirection() const { return m_ny; }
	bool GetCaps() const { return m_caps; }
	LEtype GetLEtype() const { return m_LEtype; }
protected:
	double m_R, m_C, m_L;
	int m_ny;
	bool m_caps;
	LEtype m_LEtype;
};

//! FDTD/extensions/operator_ext_excitation.cpp:166-290, operator_ext_tfsf.cpp:116-176
class CSPropExcitation : public CSProperties
{
public:
	CSPropExcitation() : CSProperties(EXCITATION), m_type(0), m_enabled(true), m_delay(0), m_freq(0)
	{
		for (int n = 0; n < 3; ++n) { m_exc[n] = 0; m_active[n] = true; m_propDir[n] = 0; }
	}
	void SetExcitType(int t) { m_type = t; }
	int GetExcitType() const { return m_type; }
	void SetEnabled(bool v) { m_enabled = v; }
	bool GetEnabled() const { return m_enabled; }
	void SetExcitation(double v, int ny) { m_exc[ny] = v; }
	double GetExcitation(int ny) const { return m_exc[ny]; }
	void SetActiveDir(bool v, int ny) { m_active[ny] = v; }
	bool GetActiveDir(int ny) const { return m_active[ny]; }
	double GetWeightedExcitation(int ny, const double*) { return m_exc[ny]; }
	void SetDelay(double d) { m_delay = d; }
	double GetDelay() const { return m_delay; }
	void SetPropagationDir(double v, int ny) { m_propDir[ny] = v; }
	double GetPropagationDir(int ny) const { return m_propDir[ny]; }
	void SetFrequency(double f) { m_freq = f; }
	double GetFrequency() const { return m_freq; }
	virtual CSPropExcitation* ToExcitation() { return this; }
protected:
	int m_type;
	bool m_enabled;
	double m_exc[3];
	bool m_active[3];
	double m_delay, m_propDir[3], m_freq;
};

//! FDTD/extensions/operator_ext_absorbing_bc.cpp:63-133
class CSPropAbsorbingBC : public CSProperties
{
public:
	enum ABCtype { UNDEFINED = 0, MUR_1ST = 1, MUR_1ST_1PV = 1, MUR_1ST_SA = 2, MUR_1ST_1PV_SA = 2 };
	CSPropAbsorbingBC() : CSProperties(ABSORBING_BC), m_normalSignPos(true), m_phaseVelocity(0), m_abcType(1) {}
	void SetNormalSignPositive(bool v) { m_normalSignPos = v; }
	bool GetNormalSignPositive() const { return m_normalSignPos; }
	void SetPhaseVelocity(double v) { m_phaseVelocity = v; }
	double GetPhaseVelocity() const { return m_phaseVelocity; }
	void SetAbsorbingBoundaryType(int t) { m_abcType = t; }
	int GetAbsorbingBoundaryType() const { return m_abcType; }
protected:
	bool m_normalSignPos;
	double m_phaseVelocity;
	int m_abcType;
};

//! Common/processing + openems.cpp:SetupProcessing only; kept so that the enum/type names resolve
class CSPropProbeBox : public CSProperties
{
public:
	CSPropProbeBox() : CSProperties(PROBEBOX) {}
};
class CSPropDumpBox : public CSPropProbeBox
{
public:
	CSPropDumpBox() : CSPropProbeBox() { m_type = DUMPBOX; }
};

//! FDTD/operator.cpp:828-833
class CSBackgroundMaterial
{
public:
	CSBackgroundMaterial() : m_epsR(1), m_mueR(1), m_kappa(0), m_sigma(0) {}
	double GetEpsilon() const { return m_epsR; }
	double GetMue() const { return m_mueR; }
	double GetKappa() const { return m_kappa; }
	double GetSigma() const { return m_sigma; }
	void SetEpsilon(double v) { m_epsR = v; }
	void SetMue(double v) { m_mueR = v; }
	void SetKappa(double v) { m_kappa = v; }
	void SetSigma(double v) { m_sigma = v; }
protected:
	double m_epsR, m_mueR, m_kappa, m_sigma;
};

class ContinuousStructure
{
public:
	ContinuousStructure() : m_nextPrimID(0) {}
	virtual ~ContinuousStructure() { for (size_t i = 0; i < m_props.size(); ++i) delete m_props[i]; }
	CSRectGrid* GetGrid() { return &m_grid; }
	CSBackgroundMaterial* GetBackgroundMaterial() { return &m_bg; }
	ParameterSet* GetParameterSet() { return &m_paraSet; }
	CoordinateSystem GetCoordInputType() const { return CARTESIAN; }
	void SetCoordInputType(CoordinateSystem) {}
	std::string Update() { return std::string(); }
	void ShowPropertyStatus(std::ostream&) {}
	void WarnUnusedPrimitves(std::ostream&) {}
	void AddProperty(CSProperties* p)
	{
		p->SetID((unsigned int)m_props.size());
		m_props.push_back(p);
	}
	//! to be called after all primitives of all properties exist (assigns IDs in creation order)
	void RegisterPrimitive(CSPrimitives* p) { p->SetID(m_nextPrimID++); m_allPrims.push_back(p); }
	size_t GetQtyProperties() const { return m_props.size(); }
	CSProperties* GetProperty(size_t n) { return m_props.at(n); }
	size_t GetQtyPropertyType(CSProperties::PropertyType type)
	{
		size_t c = 0;
		for (size_t i = 0; i < m_props.size(); ++i) if (m_props[i]->GetType() & type) ++c;
		return c;
	}
	std::vector<CSProperties*> GetPropertyByType(CSProperties::PropertyType type)
	{
		std::vector<CSProperties*> out;
		for (size_t i = 0; i < m_props.size(); ++i) if (m_props[i]->GetType() & type) out.push_back(m_props[i]);
		return out;
	}
	//! priority-sorted (descending; later primitive first on ties) list of all primitives of the given
	//! property type(s) whose bounding box intersects boundBox. FDTD/operator.cpp:1813-1838
	std::vector<CSPrimitives*> GetPrimitivesByBoundBox(const double* boundBox, bool sorted = false, CSProperties::PropertyType type = CSProperties::ANY)
	{
		std::vector<CSPrimitives*> out;
		for (size_t i = 0; i < m_allPrims.size(); ++i) {
			CSPrimitives* p = m_allPrims[i];
			if (!(p->GetProperty()->GetType() & type)) continue;
			double bb[6];
			p->GetBoundBox(bb);
			bool hit = true;
			for (int n = 0; n < 3; ++n)
				if ((bb[2*n] > boundBox[2*n+1]) || (bb[2*n+1] < boundBox[2*n])) hit = false;
			if (hit) out.push_back(p);
		}
		if (sorted) std::stable_sort(out.begin(), out.end(), higherPriority);
		return out;
	}
	std::vector<CSPrimitives*> GetAllPrimitives(bool sorted = false, CSProperties::PropertyType type = CSProperties::ANY)
	{
		std::vector<CSPrimitives*> out;
		for (size_t i = 0; i < m_allPrims.size(); ++i)
			if (m_allPrims[i]->GetProperty()->GetType() & type) out.push_back(m_allPrims[i]);
		if (sorted) std::stable_sort(out.begin(), out.end(), higherPriority);
		return out;
	}
	//! first primitive of the (priority-sorted) list that contains coord. FDTD/operator.cpp:1289,2063
	CSProperties* GetPropertyByCoordPriority(const double* coord, std::vector<CSPrimitives*> primList, bool markFoundAsUsed = false, CSPrimitives** foundPrimitive = NULL)
	{
		for (size_t i = 0; i < primList.size(); ++i) {
			if (primList[i]->IsInside(coord)) {
				if (foundPrimitive) *foundPrimitive = primList[i];
				if (markFoundAsUsed) primList[i]->SetPrimitiveUsed(true);
				return primList[i]->GetProperty();
			}
		}
		return NULL;
	}
	CSProperties* GetPropertyByCoordPriority(const double* coord, CSProperties::PropertyType type = CSProperties::ANY, bool markFoundAsUsed = false, CSPrimitives** foundPrimitive = NULL)
	{
		return GetPropertyByCoordPriority(coord, GetAllPrimitives(true, type), markFoundAsUsed, foundPrimitive);
	}
	//! NULL-terminated new[] array of all properties at coord, highest priority first
	//! (FDTD/extensions/operator_ext_absorbing_bc.cpp:211; the caller deletes the array)
	CSProperties** GetPropertiesByCoordsPriority(const double* coord, CSProperties::PropertyType type = CSProperties::ANY, bool markFoundAsUsed = false)
	{
		std::vector<CSPrimitives*> prims = GetAllPrimitives(true, type);
		std::vector<CSProperties*> found;
		for (size_t i = 0; i < prims.size(); ++i)
			if (prims[i]->IsInside(coord)) {
				if (markFoundAsUsed) prims[i]->SetPrimitiveUsed(true);
				found.push_back(prims[i]->GetProperty());
			}
		if (found.empty()) return NULL;
		CSProperties** out = new CSProperties*[found.size() + 1];
		for (size_t i = 0; i < found.size(); ++i) out[i] = found[i];
		out[found.size()] = NULL;
		return out;
	}
	bool Write2XML(const char*) { return false; }
	bool Write2XML(const std::string&) { return false; }
protected:
	static bool higherPriority(CSPrimitives* a, CSPrimitives* b)
	{
		if (a->GetPriority() != b->GetPriority()) return a->GetPriority() > b->GetPriority();
		return a->GetID() > b->GetID();
	}
	CSRectGrid m_grid;
	CSBackgroundMaterial m_bg;
	ParameterSet m_paraSet;
	std::vector<CSProperties*> m_props;
	std::vector<CSPrimitives*> m_allPrims;
	unsigned int m_nextPrimID;
};

#endif
