// TEST INFRASTRUCTURE ONLY: include path that offers the reduced frame.h WITHOUT the reduced point.h, so that the reference's own
// svo/common/point.h is the one found (libpoint_ref.so, oracle/Makefile).
#include "../../../svo_fake/svo/common/frame.h"
