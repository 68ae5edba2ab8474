// TEST INFRASTRUCTURE ONLY (oracle/): the reference's stream macros, silenced.
#pragma once
#include <iostream>
#include <sstream>
#define SVO_INFO_STREAM(x) do { std::stringstream svo_ss_; svo_ss_ << x; } while (0)
#define SVO_DEBUG_STREAM(x) do { std::stringstream svo_ss_; svo_ss_ << x; } while (0)
#define SVO_WARN_STREAM(x) do { std::stringstream svo_ss_; svo_ss_ << x; } while (0)
#define SVO_ERROR_STREAM(x) do { std::stringstream svo_ss_; svo_ss_ << x; } while (0)
#define SVO_WARN_STREAM_THROTTLE(rate, x) do { std::stringstream svo_ss_; svo_ss_ << x; } while (0)
