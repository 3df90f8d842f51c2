// Stand-in for the reference's exception.hpp (Boost.Exception): what third_party/snobal/sno.cpp uses of it.
#pragma once
#include <stdexcept>
#include <string>
struct module_error : public std::runtime_error { using std::runtime_error::runtime_error; };
#define CHM_THROW_EXCEPTION(exception_type, message) throw exception_type(std::string(message))
