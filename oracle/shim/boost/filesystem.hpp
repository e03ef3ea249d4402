// Test-infrastructure shim (NOT product code): the reference uses exactly one boost call,
// boost::filesystem::path(f).parent_path().string() (reference src/grom2atomFMO.cpp:123).
// boost is not installed in this image; std::filesystem::path has the same semantics for that call.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using path = std::filesystem::path; } }
