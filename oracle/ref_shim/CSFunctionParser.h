/* TEST INFRASTRUCTURE shim: CSXCAD's CSFunctionParser = fparser with a few predefined constants
 * (Common/processmodematch.cpp:29,123,176) */
#pragma once
#include "fparser.hh"
class CSFunctionParser : public FunctionParser
{
public:
	CSFunctionParser()
	{
		AddConstant("pi", 3.14159265358979323846);
		AddConstant("e", 2.71828182845904523536);
	}
};
