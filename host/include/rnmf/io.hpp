// io.hpp -- minimal mirror of src/io.rs write_vtk: one ASCII .vti (ImageData) file per call with
// one cell-data array per output callback (io.rs:97-198).  Off the hot path: it only needs the
// valid cells, downloaded once (get_valid_cells).
#pragma once

#include <fstream>
#include <functional>
#include <iomanip>
#include <map>
#include <string>
#include <vector>

#include "rnmf.hpp"

namespace rnmf {
namespace io {

struct VtkData {                      // io.rs:39-95: flat values + number of components
    std::vector<Real> data;
    std::size_t n_comp = 1;
};
using OutputCallBack = std::function<VtkData(const CartesianDataFrame2D&)>;
using OutputCallBackHashMap = std::map<std::string, OutputCallBack>;

// VtkData::try_from(frame): valid cells, component-interleaved per cell
inline VtkData vtk_from_frame(const CartesianDataFrame2D& f)
{
    const std::vector<Real> v = f.get_valid_cells();
    const std::size_t cells = f.underlying_mesh->n[0] * f.underlying_mesh->n[1];
    VtkData out;
    out.n_comp = f.n_comp;
    out.data.resize(v.size());
    for (std::size_t c = 0; c < f.n_comp; ++c)
        for (std::size_t p = 0; p < cells; ++p) out.data[p * f.n_comp + c] = v[c * cells + p];
    return out;
}

// write_vtk(loc, name, iter, &df, &callbacks, ..): file is `loc + "_" + name + "_" + iter + ".vti"` (io.rs:107)
inline std::string write_vtk(const std::string& loc, const std::string& name, std::size_t iter, const CartesianDataFrame2D& df,
                             const OutputCallBackHashMap& callbacks)
{
    const auto& m = *df.underlying_mesh;
    const std::string path = loc + "_" + name + "_" + std::to_string(iter) + ".vti";
    std::ofstream o(path);
    if (!o) throw Panic("write_vtk: cannot open " + path);
    o << std::setprecision(17);
    o << "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\">\n";
    o << "<ImageData WholeExtent=\"0 " << m.n[0] << " 0 " << m.n[1] << " 0 0\" Origin=\"" << m.lo[0] << " " << m.lo[1]
      << " 0\" Spacing=\"" << m.dx[0] << " " << m.dx[1] << " 1\">\n";
    o << "<Piece Extent=\"0 " << m.n[0] << " 0 " << m.n[1] << " 0 0\">\n<CellData>\n";
    for (const auto& kv : callbacks) {
        const VtkData d = kv.second(df);
        o << "<DataArray type=\"Float64\" Name=\"" << kv.first << "\" NumberOfComponents=\"" << d.n_comp << "\" format=\"ascii\">\n";
        for (std::size_t i = 0; i < d.data.size(); ++i) o << d.data[i] << ((i + 1) % 6 ? ' ' : '\n');
        o << "\n</DataArray>\n";
    }
    o << "</CellData>\n</Piece>\n</ImageData>\n</VTKFile>\n";
    return path;
}

}  // namespace io
}  // namespace rnmf
