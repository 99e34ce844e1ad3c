// circl_b200/csrc/tables.cu -- device tables built at cb200_init time.
#include "common.cuh"
#include "context.h"

namespace cb200 {
int init_extra_tables() { return 0; }
}  // namespace cb200
