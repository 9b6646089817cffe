#define VL 16
#include "sde_oracle_impl.h"
