// Stand-in for the reference's logger.hpp (spdlog is absent): logging is a no-op in the harness.
#pragma once
namespace spdlog {
template <typename... A> inline void error(A&&...) {}
template <typename... A> inline void debug(A&&...) {}
template <typename... A> inline void warn(A&&...) {}
template <typename... A> inline void info(A&&...) {}
}
#define SPDLOG_DEBUG(...) ((void)0)
#define SPDLOG_INFO(...) ((void)0)
#define SPDLOG_WARN(...) ((void)0)
#define SPDLOG_ERROR(...) ((void)0)
