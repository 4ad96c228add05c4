// Device-side launch tables of the HCA kernels.
#pragma once
#include <cstdint>
namespace cri {
struct HcaJob {};
}  // namespace cri
