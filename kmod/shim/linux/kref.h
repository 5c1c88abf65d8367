#include "../sim_kernel.h"
