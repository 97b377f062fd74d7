// Python module `spz`: the reference's Python surface (src/python/spz/spz.cc:110-361, a nanobind
// module) re-pointed at the B200-native library.  Same names, argument meaning, validation
// messages and copy-in / copy-out float32 semantics, so scripts and tests written against the
// reference module run unchanged; save_spz / load_spz reach the GPU codec through the drop-in C++
// API.  Built with pybind11 (nanobind is neither vendored in the reference checkout nor installed
// here); pybind11 words failed argument conversions the same way ("incompatible function
// arguments"), which is what the reference's tests match on (load_spz_test.py:367-372).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/spz_b200/spz.hpp"

namespace py = pybind11;

namespace {

using FloatArray = py::array_t<float, py::array::c_style | py::array::forcecast>;

// Fresh float32 1-D array owning a copy of the vector (the reference's vector_getter, :47-79).
py::array_t<float> copyOut(const std::vector<float> &v) {
  py::array_t<float> a((py::ssize_t)v.size());
  if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(float));
  return a;
}

// 1-D real numeric array -> float32 copy in `dst` (the reference's NdArray1D + vector_setter,
// :21, :91-104): ints / float64 convert, strings / complex / 2-D are a TypeError.
void copyIn(const char *name, const py::object &value, std::vector<float> *dst) {
  if (!py::isinstance<py::array>(value) && !py::isinstance<py::sequence>(value))
    throw py::type_error(std::string(name) + ": incompatible function arguments (expected a 1-D numeric array)");
  py::array raw = py::array::ensure(value);
  if (!raw) throw py::type_error(std::string(name) + ": incompatible function arguments (not array-like)");
  const char kind = raw.dtype().kind();
  if (!(kind == 'f' || kind == 'i' || kind == 'u' || kind == 'b') || raw.ndim() != 1)
    throw py::type_error(std::string(name) +
                         ": incompatible function arguments (expected a 1-D real numeric array convertible to float32)");
  FloatArray arr = FloatArray::ensure(raw);
  if (!arr) throw py::type_error(std::string(name) + ": incompatible function arguments (cannot convert to float32)");
  const size_t n = (size_t)arr.shape(0);
  dst->resize(n);
  if (n) std::memcpy(dst->data(), arr.data(), n * sizeof(float));
}

void requireMultiple(const char *name, size_t size, size_t k) {
  if (size % k != 0)
    throw py::value_error(std::string(name) + " length must be a multiple of " + std::to_string(k) + ", got " +
                          std::to_string(size));
}

}  // namespace

PYBIND11_MODULE(spz, m) {
  m.doc() = "Python bindings for the spz library (Gaussian splatting), B200-native codec.";

  py::enum_<spz::CoordinateSystem>(m, "CoordinateSystem",
                                   "Axis conventions: R/L (x right/left), U/D (y up/down), F/B (z front/back).")
      .value("UNSPECIFIED", spz::CoordinateSystem::UNSPECIFIED, "Unspecified coordinate system")
      .value("LDB", spz::CoordinateSystem::LDB, "Left Down Back")
      .value("RDB", spz::CoordinateSystem::RDB, "Right Down Back")
      .value("LUB", spz::CoordinateSystem::LUB, "Left Up Back")
      .value("RUB", spz::CoordinateSystem::RUB, "Right Up Back (Three.js)")
      .value("LDF", spz::CoordinateSystem::LDF, "Left Down Front")
      .value("RDF", spz::CoordinateSystem::RDF, "Right Down Front (PLY format)")
      .value("LUF", spz::CoordinateSystem::LUF, "Left Up Front (GLB format)")
      .value("RUF", spz::CoordinateSystem::RUF, "Right Up Front (Unity)")
      .export_values();

  py::class_<spz::PackOptions>(m, "PackOptions")
      .def(py::init<>())
      .def_readwrite("from_coord", &spz::PackOptions::from, "Coordinate system of the input splat");

  py::class_<spz::UnpackOptions>(m, "UnpackOptions")
      .def(py::init<>())
      .def_readwrite("to_coord", &spz::UnpackOptions::to, "Desired coordinate system of the output splat");

  using Cloud = spz::GaussianCloud;
  auto pointCount = [](const Cloud &c) { return (int32_t)(c.positions.size() / 3); };

  py::class_<Cloud>(m, "GaussianCloud")
      .def(py::init<>(), "Construct an empty GaussianCloud.")
      .def_property_readonly("num_points", pointCount, "Number of gaussians (derived from positions).")
      .def("__len__", pointCount)
      .def("__repr__",
           [pointCount](const Cloud &c) {
             return "GaussianCloud(num_points=" + std::to_string(pointCount(c)) + ", sh_degree=" + std::to_string(c.shDegree) +
                    ", antialiased=" + (c.antialiased ? "True" : "False") + ")";
           })
      .def_property(
          "sh_degree", [](const Cloud &c) { return c.shDegree; },
          [](Cloud &c, int32_t degree) {
            if (degree < 0 || degree > 3) throw py::value_error("sh_degree must be in [0, 3]");
            c.shDegree = degree;
          },
          "Degree of spherical harmonics (0..3).")
      .def_readwrite("antialiased", &Cloud::antialiased, "Render with mip-splatting antialiasing.")
      .def_property(
          "positions", [](const Cloud &c) { return copyOut(c.positions); },
          [](Cloud &c, const py::object &v) {
            std::vector<float> tmp;
            copyIn("positions", v, &tmp);
            requireMultiple("positions", tmp.size(), 3);
            c.positions.swap(tmp);
            c.numPoints = (int32_t)(c.positions.size() / 3);  // positions define num_points
          },
          "Gaussian centers, flat xyz; setting them defines num_points.")
      .def_property(
          "scales", [](const Cloud &c) { return copyOut(c.scales); },
          [](Cloud &c, const py::object &v) {
            std::vector<float> tmp;
            copyIn("scales", v, &tmp);
            requireMultiple("scales", tmp.size(), 3);
            c.scales.swap(tmp);
            if (c.numPoints > 0 && c.scales.size() != (size_t)c.numPoints * 3)
              throw py::value_error("scales length must equal num_points * 3");
          },
          "Log-scale radii, flat xyz.")
      .def_property(
          "rotations", [](const Cloud &c) { return copyOut(c.rotations); },
          [](Cloud &c, const py::object &v) {
            std::vector<float> tmp;
            copyIn("rotations", v, &tmp);
            requireMultiple("rotations", tmp.size(), 4);
            c.rotations.swap(tmp);
            if (c.numPoints > 0 && c.rotations.size() != (size_t)c.numPoints * 4)
              throw py::value_error("rotations length must equal num_points * 4");
          },
          "Quaternions, flat xyzw.")
      .def_property(
          "alphas", [](const Cloud &c) { return copyOut(c.alphas); },
          [](Cloud &c, const py::object &v) {
            std::vector<float> tmp;
            copyIn("alphas", v, &tmp);
            c.alphas.swap(tmp);
            if (c.numPoints > 0 && c.alphas.size() != (size_t)c.numPoints)
              throw py::value_error("alphas length must equal num_points");
          },
          "Pre-sigmoid opacities.")
      .def_property(
          "colors", [](const Cloud &c) { return copyOut(c.colors); },
          [](Cloud &c, const py::object &v) {
            std::vector<float> tmp;
            copyIn("colors", v, &tmp);
            requireMultiple("colors", tmp.size(), 3);
            c.colors.swap(tmp);
            if (c.numPoints > 0 && c.colors.size() != (size_t)c.numPoints * 3)
              throw py::value_error("colors length must equal num_points * 3");
          },
          "SH DC colour, flat rgb.")
      .def_property(
          "sh", [](const Cloud &c) { return copyOut(c.sh); },
          [](Cloud &c, const py::object &v) {
            std::vector<float> tmp;
            copyIn("sh", v, &tmp);
            requireMultiple("sh", tmp.size(), 3);
            const int d = c.shDegree;
            const size_t perChannel = d == 0 ? 0 : (size_t)((d + 1) * (d + 1) - 1);
            if (perChannel == 0) {
              if (!tmp.empty()) throw py::value_error("sh must be empty when sh_degree == 0");
            } else {
              requireMultiple("sh", tmp.size(), perChannel * 3);
            }
            c.sh.swap(tmp);
            if (c.numPoints > 0 && c.sh.size() != (size_t)c.numPoints * perChannel * 3)
              throw py::value_error("sh length must equal num_points * ((sh_degree+1)^2 - 1) * 3");
          },
          "SH coefficients, coefficient-major with rgb innermost; set sh_degree first.")
      .def("convert_coordinates", &Cloud::convertCoordinates, py::arg("from_coord"), py::arg("to_coord"),
           "Convert between two coordinate systems in place.")
      .def("rotate_180_deg_about_x", &Cloud::rotate180DegAboutX, "RUB <-> RDF (180 degrees about X).")
      .def("median_volume", &Cloud::medianVolume, "Median gaussian volume.");

  m.def("load_spz", (spz::GaussianCloud(*)(const std::string &, const spz::UnpackOptions &)) & spz::loadSpz,
        py::arg("filename"), py::arg("options") = spz::UnpackOptions(), py::call_guard<py::gil_scoped_release>(),
        "Load a *.spz* file and return a GaussianCloud.");
  m.def("save_spz", (bool (*)(const spz::GaussianCloud &, const spz::PackOptions &, const std::string &)) & spz::saveSpz,
        py::arg("gaussians"), py::arg("options"), py::arg("filename"), py::call_guard<py::gil_scoped_release>(),
        "Save a GaussianCloud to a *.spz* file.");
  m.def("load_splat_from_ply", &spz::loadSplatFromPly, py::arg("filename"), py::arg("options") = spz::UnpackOptions(),
        py::call_guard<py::gil_scoped_release>(), "Load a .ply splat file.");
  m.def("save_splat_to_ply", &spz::saveSplatToPly, py::arg("gaussians"), py::arg("options"), py::arg("filename"),
        py::call_guard<py::gil_scoped_release>(), "Save a GaussianCloud as a .ply splat file.");
}
