// Minimal stand-in for <rclcpp/rclcpp.hpp> (TEST INFRASTRUCTURE ONLY): the hot-path translation units only use the
// logger from ORB_SLAM2/Error.h:19.
#pragma once
#include <cstdio>
namespace rclcpp
{
struct Logger
{
  const char *name;
};
static inline Logger get_logger(const char *name) { return Logger{name}; }
} // namespace rclcpp
#define RCLCPP_ERROR(logger, ...)                      \
  do                                                   \
  {                                                    \
    std::fprintf(stderr, "[ERROR] [%s]: ", (logger).name); \
    std::fprintf(stderr, __VA_ARGS__);                 \
    std::fprintf(stderr, "\n");                        \
  } while (0)
#define RCLCPP_INFO(logger, ...) \
  do                             \
  {                              \
  } while (0)
