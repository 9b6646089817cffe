#define VL 8
#include "sde_oracle_impl.h"
