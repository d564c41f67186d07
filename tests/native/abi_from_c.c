/* TEST: include/vh_femgl.h is a plain C header (C99, -pedantic) and the library links from C: the drop-in boundary has no
 * C++ or torch types in it.  Calls only the host-side validator, so it runs without a GPU. */
#include "vh_femgl.h"

#include <stdio.h>
#include <string.h>

int main(void)
{
  vh_mesh_desc d;
  char         msg[128];
  int32_t      cell_nodes[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  double       origin[3] = {0, 0, 0}, h[3] = {1, 1, 1};
  memset(&d, 0, sizeof d);
  d.degree = 3;
  if (vh_validate_mesh_desc(&d, msg, (int)sizeof msg) != VH_ERR_ARG || !strstr(msg, "degree"))
    return 1;
  d.degree        = 1;
  d.n_owned_nodes = 8;
  d.n_cells       = 1;
  d.cell_nodes    = cell_nodes;
  d.cell_origin   = origin;
  d.cell_h        = h;
  if (vh_validate_mesh_desc(&d, msg, (int)sizeof msg) != VH_OK)
    return 2;
  cell_nodes[7] = 8;
  if (vh_validate_mesh_desc(&d, msg, (int)sizeof msg) != VH_ERR_ARG)
    return 3;
  if (vh_last_error(NULL) == NULL)
    return 4;
  printf("ok\n");
  return 0;
}
