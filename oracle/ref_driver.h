/* TEST INFRASTRUCTURE (oracle/_ref build only). C API of the reference harness: the same calls as
 * oracle/fdtd_oracle.h with the prefix ref_ (served by the reference's own classes), plus a few extras
 * (engine choice, Processing classes, recorded dumps). See ref_driver.cpp. */
#pragma once
struct ref_sim;
