// Test-infrastructure shim (NOT product code): the reference includes <libconfig.h++>
// (src/grom2atomFMO.cpp:6) but only uses it under #ifdef OLDPARSER, which its build never defines.
#pragma once
