/* RegularGrid.h -- same file name as the reference's header (AnisotropicElastoplasticity/RegularGrid.h): put include/aep/compat on the include
 * path in place of the reference's source directory and `#include "RegularGrid.h"` resolves to the B200 host class. */
#include "../RegularGrid.h"
