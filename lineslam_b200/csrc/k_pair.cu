#include "lsl_internal.h"
