// abi_internal.h — shared by the translation units that implement include/wepp_b200.h
#pragma once
#include <string>

namespace wepp {
int abi_fail(int code, const std::string& msg);   // sets wepp_last_error() and returns code
}
