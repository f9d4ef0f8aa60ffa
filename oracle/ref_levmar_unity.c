/* oracle/_ref build recipe for levmar 2.6 (TEST INFRASTRUCTURE).
 * Compiles the reference's own sources where they lie (-I$(REF)/external/levmar-2.6);
 * nothing is copied. levmar.h:31 hard-defines HAVE_LAPACK, and no LAPACK exists in this
 * image, so the header is included once here, the macro dropped, and the three
 * translation units the dlevmar_dif path needs are pulled in (their own
 * #include "levmar.h" is then a no-op thanks to the include guard). The built-in
 * LU of Axb_core.c:1140 replaces dgetrf/dgetrs (Tier-T difference). */
#include "levmar.h"
#undef HAVE_LAPACK
#include "lm.c"
#include "misc.c"
#include "Axb.c"
