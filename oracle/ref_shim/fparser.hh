/* TEST INFRASTRUCTURE (oracle/_ref build only). Stand-in for the fparser library (thliebig/fparser, master,
 * unpinned; not vendored under /root/reference): class FunctionParser with the calls the reference makes
 * (FDTD/extensions/operator_ext_upml.cpp:28,261-333, FDTD/excitation.cpp:231-246,
 * Common/processmodematch.cpp:123,176). A recursive-descent evaluator over doubles:
 * + - * / % ^, unary -, comparison, parentheses, constants, variables and the usual <cmath> functions,
 * evaluated strictly left to right as written (no algebraic re-association).
 */
#ifndef FPARSER_SHIM_HH
#define FPARSER_SHIM_HH
#include <string>
#include <vector>
#include <map>
#include <cmath>
#include <cstdlib>
#include <cctype>

class FunctionParser
{
public:
	enum ParseErrorType { SYNTAX_ERROR = 0, MISM_PARENTH, MISSING_PARENTH, EMPTY_PARENTH, EXPECT_OPERATOR, OUT_OF_MEMORY,
	                      UNEXPECTED_ERROR, INVALID_VARS, ILL_PARAMS_AMOUNT, PREMATURE_EOS, EXPECT_PARENTH_FUNC,
	                      UNKNOWN_IDENTIFIER, NO_FUNCTION_PARSED_YET, FP_NO_ERROR };
	FunctionParser() : m_err(NO_FUNCTION_PARSED_YET), m_evalErr(0) {}
	virtual ~FunctionParser() {}
	bool AddConstant(const std::string& name, double v) { m_const[name] = v; return true; }
	//! returns -1 on success, else the position of the error
	int Parse(const std::string& func, const std::string& vars, bool = false)
	{
		m_vars.clear();
		std::string cur;
		for (size_t i = 0; i <= vars.size(); ++i) {
			if (i == vars.size() || vars[i] == ',') { if (!cur.empty()) m_vars.push_back(cur); cur.clear(); }
			else if (!isspace((unsigned char)vars[i])) cur.push_back(vars[i]);
		}
		m_code.clear();
		m_src = func;
		m_pos = 0;
		m_err = FP_NO_ERROR;
		parseExpr();
		skip();
		if (m_err == FP_NO_ERROR && m_pos != m_src.size()) m_err = SYNTAX_ERROR;
		if (m_err != FP_NO_ERROR) { m_code.clear(); return (int)m_pos; }
		return -1;
	}
	int Parse(const char* func, const std::string& vars, bool d = false) { return Parse(std::string(func), vars, d); }
	ParseErrorType GetParseErrorType() const { return m_err; }
	const char* ErrorMsg() const { return m_err == FP_NO_ERROR ? "" : "fparser shim: parse error"; }
	int EvalError() const { return m_evalErr; }
	void Optimize() {}
	double Eval(const double* vars)
	{
		m_evalErr = 0;
		std::vector<double> st;
		st.reserve(16);
		for (size_t i = 0; i < m_code.size(); ++i) {
			const Op& o = m_code[i];
			switch (o.kind) {
			case K_NUM: st.push_back(o.val); break;
			case K_VAR: st.push_back(vars[o.idx]); break;
			case K_NEG: st.back() = -st.back(); break;
			case K_NOT: st.back() = (fabs(st.back()) < 0.5) ? 1.0 : 0.0; break;
			case K_BIN: { double b = st.back(); st.pop_back(); double a = st.back(); st.back() = bin(o.idx, a, b); break; }
			case K_FUN: {
				int na = o.nargs;
				double a[3] = {0, 0, 0};
				for (int k = na - 1; k >= 0; --k) { a[k] = st.back(); st.pop_back(); }
				st.push_back(fun(o.idx, a));
				break; }
			}
		}
		return st.empty() ? 0.0 : st.back();
	}
private:
	enum Kind { K_NUM, K_VAR, K_NEG, K_NOT, K_BIN, K_FUN };
	struct Op { Kind kind; int idx; int nargs; double val; };
	enum { B_ADD, B_SUB, B_MUL, B_DIV, B_MOD, B_POW, B_LT, B_LE, B_GT, B_GE, B_EQ, B_NE, B_AND, B_OR };
	enum { F_SIN, F_COS, F_TAN, F_ASIN, F_ACOS, F_ATAN, F_ATAN2, F_SINH, F_COSH, F_TANH, F_EXP, F_LOG, F_LOG10, F_LOG2, F_SQRT,
	       F_POW, F_ABS, F_MIN, F_MAX, F_FLOOR, F_CEIL, F_INT, F_IF, F_EXP2, F_CBRT, F_HYPOT, F_TRUNC, F_COUNT };
	static double bin(int op, double a, double b)
	{
		switch (op) {
		case B_ADD: return a + b; case B_SUB: return a - b; case B_MUL: return a * b; case B_DIV: return a / b;
		case B_MOD: return fmod(a, b); case B_POW: return pow(a, b);
		case B_LT: return a < b; case B_LE: return a <= b; case B_GT: return a > b; case B_GE: return a >= b;
		case B_EQ: return a == b; case B_NE: return a != b;
		case B_AND: return (fabs(a) >= 0.5) && (fabs(b) >= 0.5); case B_OR: return (fabs(a) >= 0.5) || (fabs(b) >= 0.5);
		}
		return 0;
	}
	static double fun(int f, const double* a)
	{
		switch (f) {
		case F_SIN: return sin(a[0]); case F_COS: return cos(a[0]); case F_TAN: return tan(a[0]);
		case F_ASIN: return asin(a[0]); case F_ACOS: return acos(a[0]); case F_ATAN: return atan(a[0]);
		case F_ATAN2: return atan2(a[0], a[1]); case F_SINH: return sinh(a[0]); case F_COSH: return cosh(a[0]);
		case F_TANH: return tanh(a[0]); case F_EXP: return exp(a[0]); case F_LOG: return log(a[0]);
		case F_LOG10: return log10(a[0]); case F_LOG2: return log2(a[0]); case F_SQRT: return sqrt(a[0]);
		case F_POW: return pow(a[0], a[1]); case F_ABS: return fabs(a[0]); case F_MIN: return a[0] < a[1] ? a[0] : a[1];
		case F_MAX: return a[0] > a[1] ? a[0] : a[1]; case F_FLOOR: return floor(a[0]); case F_CEIL: return ceil(a[0]);
		case F_INT: return floor(a[0] + 0.5); case F_IF: return (fabs(a[0]) >= 0.5) ? a[1] : a[2];
		case F_EXP2: return exp2(a[0]); case F_CBRT: return cbrt(a[0]); case F_HYPOT: return hypot(a[0], a[1]);
		case F_TRUNC: return trunc(a[0]);
		}
		return 0;
	}
	void skip() { while (m_pos < m_src.size() && isspace((unsigned char)m_src[m_pos])) ++m_pos; }
	bool eat(const char* tok)
	{
		skip();
		size_t n = strlen_(tok);
		if (m_src.compare(m_pos, n, tok) == 0) { m_pos += n; return true; }
		return false;
	}
	static size_t strlen_(const char* s) { size_t n = 0; while (s[n]) ++n; return n; }
	void emit(Kind k, int idx = 0, int nargs = 0, double v = 0) { Op o; o.kind = k; o.idx = idx; o.nargs = nargs; o.val = v; m_code.push_back(o); }
	void fail(ParseErrorType e) { if (m_err == FP_NO_ERROR) m_err = e; }
	// precedence (low to high): | & comparison +- */% unary ^
	void parseExpr() { parseAnd(); while (m_err == FP_NO_ERROR && eat("|")) { parseAnd(); emit(K_BIN, B_OR); } }
	void parseAnd() { parseCmp(); while (m_err == FP_NO_ERROR && eat("&")) { parseCmp(); emit(K_BIN, B_AND); } }
	void parseCmp()
	{
		parseAdd();
		while (m_err == FP_NO_ERROR) {
			int op;
			if (eat("<=")) op = B_LE; else if (eat(">=")) op = B_GE; else if (eat("!=")) op = B_NE;
			else if (eat("<")) op = B_LT; else if (eat(">")) op = B_GT; else if (eat("=")) op = B_EQ; else break;
			parseAdd();
			emit(K_BIN, op);
		}
	}
	void parseAdd()
	{
		parseMul();
		while (m_err == FP_NO_ERROR) {
			int op;
			if (eat("+")) op = B_ADD; else if (eat("-")) op = B_SUB; else break;
			parseMul();
			emit(K_BIN, op);
		}
	}
	void parseMul()
	{
		parseUnary();
		while (m_err == FP_NO_ERROR) {
			int op;
			if (eat("*")) op = B_MUL; else if (eat("/")) op = B_DIV; else if (eat("%")) op = B_MOD; else break;
			parseUnary();
			emit(K_BIN, op);
		}
	}
	void parseUnary()
	{
		if (eat("-")) { parseUnary(); emit(K_NEG); return; }
		if (eat("+")) { parseUnary(); return; }
		if (eat("!")) { parseUnary(); emit(K_NOT); return; }
		parsePow();
	}
	void parsePow()
	{
		parseAtom();
		if (m_err == FP_NO_ERROR && eat("^")) { parseUnary(); emit(K_BIN, B_POW); }   // right associative, binds tighter than unary minus on its left
	}
	void parseAtom()
	{
		skip();
		if (m_pos >= m_src.size()) { fail(PREMATURE_EOS); return; }
		char c = m_src[m_pos];
		if (c == '(') {
			++m_pos;
			parseExpr();
			if (!eat(")")) fail(MISSING_PARENTH);
			return;
		}
		if (isdigit((unsigned char)c) || c == '.') {
			const char* s = m_src.c_str() + m_pos;
			char* end = 0;
			double v = strtod(s, &end);
			if (end == s) { fail(SYNTAX_ERROR); return; }
			m_pos += (size_t)(end - s);
			emit(K_NUM, 0, 0, v);
			return;
		}
		if (isalpha((unsigned char)c) || c == '_') {
			size_t b = m_pos;
			while (m_pos < m_src.size() && (isalnum((unsigned char)m_src[m_pos]) || m_src[m_pos] == '_')) ++m_pos;
			std::string id = m_src.substr(b, m_pos - b);
			for (size_t i = 0; i < m_vars.size(); ++i) if (m_vars[i] == id) { emit(K_VAR, (int)i); return; }
			std::map<std::string, double>::const_iterator it = m_const.find(id);
			if (it != m_const.end()) { emit(K_NUM, 0, 0, it->second); return; }
			static const struct { const char* name; int id; int nargs; } funcs[] = {
				{"sin", F_SIN, 1}, {"cos", F_COS, 1}, {"tan", F_TAN, 1}, {"asin", F_ASIN, 1}, {"acos", F_ACOS, 1}, {"atan", F_ATAN, 1},
				{"atan2", F_ATAN2, 2}, {"sinh", F_SINH, 1}, {"cosh", F_COSH, 1}, {"tanh", F_TANH, 1}, {"exp", F_EXP, 1}, {"log", F_LOG, 1},
				{"log10", F_LOG10, 1}, {"log2", F_LOG2, 1}, {"sqrt", F_SQRT, 1}, {"pow", F_POW, 2}, {"abs", F_ABS, 1}, {"min", F_MIN, 2},
				{"max", F_MAX, 2}, {"floor", F_FLOOR, 1}, {"ceil", F_CEIL, 1}, {"int", F_INT, 1}, {"if", F_IF, 3}, {"exp2", F_EXP2, 1},
				{"cbrt", F_CBRT, 1}, {"hypot", F_HYPOT, 2}, {"trunc", F_TRUNC, 1} };
			for (size_t i = 0; i < sizeof(funcs) / sizeof(funcs[0]); ++i) {
				if (id == funcs[i].name) {
					if (!eat("(")) { fail(EXPECT_PARENTH_FUNC); return; }
					for (int a = 0; a < funcs[i].nargs; ++a) {
						if (a && !eat(",")) { fail(ILL_PARAMS_AMOUNT); return; }
						parseExpr();
					}
					if (!eat(")")) { fail(MISSING_PARENTH); return; }
					emit(K_FUN, funcs[i].id, funcs[i].nargs);
					return;
				}
			}
			m_pos = b;
			fail(UNKNOWN_IDENTIFIER);
			return;
		}
		fail(SYNTAX_ERROR);
	}
	std::vector<std::string> m_vars;
	std::map<std::string, double> m_const;
	std::vector<Op> m_code;
	std::string m_src;
	size_t m_pos;
	ParseErrorType m_err;
	int m_evalErr;
};
#endif
