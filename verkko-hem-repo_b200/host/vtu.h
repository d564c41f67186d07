// VTU / PVTU writer of the host mirror (output_results, /root/reference/femgl/src/io.cc:106-170); see vtu.cc.
#ifndef VH_HOST_VTU_H
#define VH_HOST_VTU_H

#include "mesh.h"

#include <string>

namespace vhhost
{
// Writes <dir>/<basename>_<counter>.<rank>.vtu (and the .pvtu record on rank 0) and returns the path of the piece.
// solution_local / update_local: 18 values per LOCAL node (owned first, then ghosts) — what vh_snapshot_wait hands out;
// a null pointer writes zeros for that block.  Throws std::runtime_error on I/O errors.
std::string write_vtu_piece(const RankTables &T, int rank, int n_ranks, const std::string &dir, const std::string &basename, int counter,
                            const double *solution_local, const double *update_local);
} // namespace vhhost
#endif
