// Stand-in for the reference's logger.hpp (spdlog/Boost.Log are absent): logging is a no-op in the harness.
#pragma once
#include <iostream>
#define SPDLOG_DEBUG(...) ((void)0)
#define SPDLOG_INFO(...) ((void)0)
#define SPDLOG_WARN(...) ((void)0)
#define SPDLOG_ERROR(...) ((void)0)
#define LOG_DEBUG if (false) std::cerr
