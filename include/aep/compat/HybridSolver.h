/* HybridSolver.h -- same file name as the reference's header (AnisotropicElastoplasticity/HybridSolver.h): put include/aep/compat on the include
 * path in place of the reference's source directory and `#include "HybridSolver.h"` resolves to the B200 host class. */
#include "../HybridSolver.h"
