/* TEST INFRASTRUCTURE shim: openems.cpp:48 prints BOOST_LIB_VERSION */
#pragma once
#define BOOST_LIB_VERSION "shim"
#define BOOST_VERSION 0
