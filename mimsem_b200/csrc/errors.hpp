// Thread-local last-error text behind mimsem_last_error() (include/mimsem_gpu.h).
#pragma once
#include <string>

namespace mimsem {
void set_error(const std::string& msg);
const char* get_error();
}  // namespace mimsem
