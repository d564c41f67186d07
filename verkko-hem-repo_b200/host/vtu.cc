// Output path of the host mirror: DataOut::write_vtu_with_pvtu_record stand-in
//   FemGL::output_results()   /root/reference/femgl/src/io.cc:106-170   (called after every Newton step, run.cc:221-227)
// One .vtu piece per rank with the rank's OWNED cells over its local (owned + ghost) nodes, the 18 components of the Newton
// update (du_11 .. dv_33, io.cc:110-128) and of the solution (u_11 .. v_33, io.cc:130-148) as point data and the subdomain
// id as cell data (io.cc:158-161); rank 0 also writes the .pvtu record.  DataOut::build_patches() with its default of one
// subdivision writes the 8 vertex values of every cell, so Q2 cells are written as linear hexahedra over their vertices.
// File names follow deal.II: <dir>/solution_<counter, 2 digits>.<rank>.vtu and <dir>/solution_<counter>.pvtu.
// Format: VTK XML UnstructuredGrid, appended raw binary (header_type UInt64): the file is written with a handful of
// fwrite calls of whole arrays, which is what keeps the writer thread far below a Newton step at GPU speeds.
#include "vtu.h"

#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace vhhost
{
namespace
{
const char *const comp_suffix[9] = {"11", "12", "13", "21", "22", "23", "31", "32", "33"};

std::string field_name(int block, int c)
{ // block 0: Newton update, 1: solution; c < 9: real part u, c >= 9: imaginary part v (femgl.cc:133-137)
  std::string s = block == 0 ? "d" : "";
  s += c < 9 ? "u_" : "v_";
  s += comp_suffix[c % 9];
  return s;
}

void make_dir(const std::string &dir)
{
  if (dir.empty())
    return;
  std::string d = dir;
  while (d.size() > 1 && d.back() == '/')
    d.pop_back();
  if (mkdir(d.c_str(), 0777) != 0)
    {
      struct stat st;
      if (stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode))
        throw std::runtime_error("write_vtu: cannot create directory " + d);
    }
}

std::string piece_name(const std::string &basename, int counter, int rank)
{
  char buf[96];
  std::snprintf(buf, sizeof buf, "%s_%02d.%d.vtu", basename.c_str(), counter, rank);
  return buf;
}
} // namespace

std::string write_vtu_piece(const RankTables &T, int rank, int n_ranks, const std::string &dir, const std::string &basename, int counter,
                            const double *solution_local, const double *update_local)
{
  const int     nn = T.degree == 1 ? 8 : 27;
  const int64_t n_points = (int64_t)T.n_owned_nodes + T.n_ghost_nodes;
  int64_t       n_cells = 0;
  for (int32_t e = 0; e < T.n_cells; ++e)
    n_cells += T.cell_owned[e] ? 1 : 0;
  make_dir(dir);
  const std::string path = dir + piece_name(basename, counter, rank);
  FILE             *f = std::fopen(path.c_str(), "wb");
  if (!f)
    throw std::runtime_error("write_vtu: cannot open " + path);

  // appended-data layout: every array is [uint64 byte count][raw bytes]
  uint64_t off = 0;
  auto     next = [&](uint64_t bytes) {
    const uint64_t o = off;
    off += 8 + bytes;
    return o;
  };
  std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n");
  std::fprintf(f, "<UnstructuredGrid>\n<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n", (long long)n_points, (long long)n_cells);
  std::fprintf(f, "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"appended\" offset=\"%llu\"/>\n</Points>\n",
               (unsigned long long)next(8ull * 3 * n_points));
  std::fprintf(f, "<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"appended\" offset=\"%llu\"/>\n",
               (unsigned long long)next(8ull * 8 * n_cells));
  std::fprintf(f, "<DataArray type=\"Int64\" Name=\"offsets\" format=\"appended\" offset=\"%llu\"/>\n", (unsigned long long)next(8ull * n_cells));
  std::fprintf(f, "<DataArray type=\"UInt8\" Name=\"types\" format=\"appended\" offset=\"%llu\"/>\n</Cells>\n", (unsigned long long)next(1ull * n_cells));
  std::fprintf(f, "<PointData Scalars=\"scalars\">\n");
  for (int block = 0; block < 2; ++block)
    for (int c = 0; c < 18; ++c)
      std::fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" format=\"appended\" offset=\"%llu\"/>\n", field_name(block, c).c_str(),
                   (unsigned long long)next(8ull * n_points));
  std::fprintf(f, "</PointData>\n<CellData>\n<DataArray type=\"Float32\" Name=\"subdomain\" format=\"appended\" offset=\"%llu\"/>\n</CellData>\n",
               (unsigned long long)next(4ull * n_cells));
  std::fprintf(f, "</Piece>\n</UnstructuredGrid>\n<AppendedData encoding=\"raw\">\n_");

  auto put = [&](const void *data, uint64_t bytes) {
    std::fwrite(&bytes, 8, 1, f);
    if (bytes)
      std::fwrite(data, 1, bytes, f);
  };
  put(T.node_xyz.data(), 8ull * 3 * n_points);
  { // VTK_HEXAHEDRON orders the vertices counter-clockwise per z layer: lexicographic 0 1 3 2 | 4 5 7 6
    static const int     perm[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    std::vector<int64_t> conn((size_t)8 * n_cells), offs((size_t)n_cells);
    std::vector<uint8_t> types((size_t)n_cells, 12);
    int64_t              k = 0;
    for (int32_t e = 0; e < T.n_cells; ++e)
      if (T.cell_owned[e])
        {
          for (int v = 0; v < 8; ++v)
            conn[(size_t)8 * k + v] = T.cell_nodes[(size_t)e * nn + perm[v]];
          offs[(size_t)k] = 8 * (k + 1);
          ++k;
        }
    put(conn.data(), 8ull * 8 * n_cells);
    put(offs.data(), 8ull * n_cells);
    put(types.data(), 1ull * n_cells);
  }
  std::vector<double> col((size_t)n_points);
  for (int block = 0; block < 2; ++block)
    {
      const double *src = block == 0 ? update_local : solution_local;
      for (int c = 0; c < 18; ++c)
        {
          for (int64_t i = 0; i < n_points; ++i)
            col[(size_t)i] = src ? src[18 * i + c] : 0.0;
          put(col.data(), 8ull * n_points);
        }
    }
  {
    std::vector<float> sub((size_t)n_cells, (float)rank);
    put(sub.data(), 4ull * n_cells);
  }
  std::fprintf(f, "\n</AppendedData>\n</VTKFile>\n");
  if (std::fclose(f) != 0)
    throw std::runtime_error("write_vtu: write error on " + path);

  if (rank == 0)
    {
      char name[64];
      std::snprintf(name, sizeof name, "%s_%02d.pvtu", basename.c_str(), counter);
      FILE *p = std::fopen((dir + name).c_str(), "wb");
      if (!p)
        throw std::runtime_error("write_vtu: cannot open " + dir + name);
      std::fprintf(p, "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n");
      std::fprintf(p, "<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n<PPointData Scalars=\"scalars\">\n");
      for (int block = 0; block < 2; ++block)
        for (int c = 0; c < 18; ++c)
          std::fprintf(p, "<PDataArray type=\"Float64\" Name=\"%s\"/>\n", field_name(block, c).c_str());
      std::fprintf(p, "</PPointData>\n<PCellData>\n<PDataArray type=\"Float32\" Name=\"subdomain\"/>\n</PCellData>\n");
      for (int r = 0; r < n_ranks; ++r)
        std::fprintf(p, "<Piece Source=\"%s\"/>\n", piece_name(basename, counter, r).c_str());
      std::fprintf(p, "</PUnstructuredGrid>\n</VTKFile>\n");
      std::fclose(p);
    }
  return path;
}
} // namespace vhhost
