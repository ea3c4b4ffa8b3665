/* test-infrastructure shim (oracle/_ref build): stands in for <boost/predef.h>, which the reference
 * includes only for BOOST_ARCH_X86 (tools/denormal.h:1-5, FDTD/engine_sse_compressed.cpp:21) */
#pragma once
#if defined(__x86_64__) || defined(__i386__)
#define BOOST_ARCH_X86 1
#else
#define BOOST_ARCH_X86 0
#endif
