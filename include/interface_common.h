/* Forwarding header: keeps the reference include line "kspace-neutrinos/interface_common.h" working.
 * All declarations live in kspace_neutrinos.h (each cites the reference line it replaces). */
#include "kspace_neutrinos.h"
