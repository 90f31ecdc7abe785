// host_common.h -- error / launch-count plumbing shared by the translation units of liblasr_b200.
#pragma once
#include <string>

namespace lasr {
int lasr_fail(const std::string& msg);      // records the message for lasr_b200_last_error(), returns 1
int lasr_check_launch(const char* what);    // counts the launch, turns cudaGetLastError() into a return code
}  // namespace lasr
