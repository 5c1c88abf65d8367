/* glibc pulls the real UAPI header of this name through <errno.h>/<sys/ioctl.h>: chain to it first */
#include_next <linux/types.h>
#include "../sim_kernel.h"
