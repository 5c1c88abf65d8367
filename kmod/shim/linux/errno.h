/* glibc pulls the real UAPI header of this name through <errno.h>/<sys/ioctl.h>: chain to it first */
#include_next <linux/errno.h>
#include "../sim_kernel.h"
