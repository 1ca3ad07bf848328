/* Forwarding header: keeps the reference include line "kspace-neutrinos/delta_tot_table.h" working.
 * All declarations live in kspace_neutrinos.h (each cites the reference line it replaces). */
#include "kspace_neutrinos.h"
