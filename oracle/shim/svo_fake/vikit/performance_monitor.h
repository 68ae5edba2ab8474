// TEST INFRASTRUCTURE ONLY (oracle/): timing/logging helper of the reference, not part of the arithmetic.
#pragma once
namespace vk { class PerformanceMonitor {}; }
