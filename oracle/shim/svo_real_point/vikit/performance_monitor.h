// TEST INFRASTRUCTURE ONLY: see svo/common/frame.h in this directory.
#include "../../svo_fake/vikit/performance_monitor.h"
