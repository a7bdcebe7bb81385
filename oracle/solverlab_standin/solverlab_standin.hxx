/* solverlab_standin.hxx -- TEST INFRASTRUCTURE ONLY.  A minimal stand-in for the SOLVERLAB (CDMATH) classes that the
 * reference's src/TransportEquation.cxx and src/WaveSystem.cxx use -- Mesh, Cell, Face, Vector, Matrix, Field,
 * CdmathException -- so that THOSE FILES, unmodified and compiled from where they lie under /root/reference, can assemble
 * their matrices in this image (no SOLVERLAB / MEDCoupling here): oracle/Makefile target _ref/libreference_assembly.so,
 * used by tests/test_reference_assembly.py only.
 *
 * What is the reference's and what is ours: computeDivergenceMatrix of both files (the upwind choice and its signs,
 * jacobianMatrices, the interior / wall / periodic / Neumann cases, the block placement through addValue) runs as
 * written.  The stand-in holds a mesh as plain finite-volume connectivity handed in by the caller (cells -> faces with
 * outward unit normals, face measures, the one or two cells of a face, border group names, periodic twins, cell measures
 * and centres): geometry in, the reference's assembly out.  Only the members those two files call exist.
 */
#ifndef SOLVERLAB_STANDIN_HXX
#define SOLVERLAB_STANDIN_HXX
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

class CdmathException : public std::runtime_error {
public:
    explicit CdmathException(const std::string &what) : std::runtime_error(what) {}
};

class Matrix {
public:
    explicit Matrix(int n = 0) : _r(n), _c(n), _v((size_t)n * n, 0.0) {}
    Matrix(int r, int c) : _r(r), _c(c), _v((size_t)r * c, 0.0) {}
    int getNumberOfRows() const { return _r; }
    int getNumberOfColumns() const { return _c; }
    double &operator()(int i, int j) { return _v[(size_t)i * _c + j]; }
    double operator()(int i, int j) const { return _v[(size_t)i * _c + j]; }
    Matrix operator-(const Matrix &o) const
    {
        if (o._r != _r || o._c != _c) throw CdmathException("Matrix - Matrix: sizes differ");
        Matrix m(_r, _c);
        for (size_t k = 0; k < _v.size(); ++k) m._v[k] = _v[k] - o._v[k];
        return m;
    }
    Matrix operator*(double s) const
    {
        Matrix m(_r, _c);
        for (size_t k = 0; k < _v.size(); ++k) m._v[k] = _v[k] * s;
        return m;
    }
    Matrix operator*(const Matrix &o) const
    {
        if (_c != o._r) throw CdmathException("Matrix * Matrix: sizes differ");
        Matrix m(_r, o._c);
        for (int i = 0; i < _r; ++i)
            for (int j = 0; j < o._c; ++j) {
                double s = 0.0;
                for (int k = 0; k < _c; ++k) s += (*this)(i, k) * o(k, j);
                m(i, j) = s;
            }
        return m;
    }

private:
    int _r, _c;
    std::vector<double> _v;
};

class Vector {
public:
    explicit Vector(int n = 0) : _v((size_t)n, 0.0) {}
    int size() const { return (int)_v.size(); }
    double &operator[](int i) { return _v[(size_t)i]; }
    double operator[](int i) const { return _v[(size_t)i]; }
    double operator*(const Vector &o) const          /* scalar product */
    {
        if (o.size() != size()) throw CdmathException("Vector * Vector: sizes differ");
        double s = 0.0;
        for (int i = 0; i < size(); ++i) s += _v[(size_t)i] * o._v[(size_t)i];
        return s;
    }
    Matrix tensProduct(const Vector &o) const        /* (u (x) v)(i, j) = u_i v_j */
    {
        Matrix m(size(), o.size());
        for (int i = 0; i < size(); ++i)
            for (int j = 0; j < o.size(); ++j) m(i, j) = _v[(size_t)i] * o._v[(size_t)j];
        return m;
    }

private:
    std::vector<double> _v;
};

struct StandinMeshData {
    int dim = 3;
    std::vector<std::vector<int>> cell_faces;            /* per cell: its face numbers */
    std::vector<std::vector<double>> cell_normals;       /* per cell: outward unit normals, dim numbers per face */
    std::vector<double> cell_measure;
    std::vector<double> cell_centre;                     /* 3 per cell */
    std::vector<double> face_measure;
    std::vector<int> face_cells;                         /* 2 per face; second = -1 on the border */
    std::vector<std::string> face_group;                 /* border faces: "Neumann", "Periodic", "Wall", ... */
    std::vector<int> face_twin;                          /* periodic twin of a border face, else -1 */
    double lo[3] = { 0, 0, 0 }, hi[3] = { 1, 1, 1 };
};

class Cell {
public:
    Cell(std::shared_ptr<const StandinMeshData> d, int j) : _d(std::move(d)), _j(j) {}
    int getNumberOfFaces() const { return (int)_d->cell_faces[(size_t)_j].size(); }
    const std::vector<int> &getFacesId() const { return _d->cell_faces[(size_t)_j]; }
    double getNormalVector(int k, int i) const { return _d->cell_normals[(size_t)_j][(size_t)k * _d->dim + i]; }
    double getMeasure() const { return _d->cell_measure[(size_t)_j]; }
    double x() const { return _d->cell_centre[3 * (size_t)_j]; }
    double y() const { return _d->cell_centre[3 * (size_t)_j + 1]; }
    double z() const { return _d->cell_centre[3 * (size_t)_j + 2]; }

private:
    std::shared_ptr<const StandinMeshData> _d;
    int _j;
};

class Face {
public:
    Face(std::shared_ptr<const StandinMeshData> d, int f) : _d(std::move(d)), _f(f) {}
    double getMeasure() const { return _d->face_measure[(size_t)_f]; }
    bool isBorder() const { return _d->face_cells[2 * (size_t)_f + 1] < 0; }
    std::vector<int> getCellsId() const
    {
        std::vector<int> c(1, _d->face_cells[2 * (size_t)_f]);
        if (!isBorder()) c.push_back(_d->face_cells[2 * (size_t)_f + 1]);
        return c;
    }
    std::string getGroupName() const { return _d->face_group[(size_t)_f]; }

private:
    std::shared_ptr<const StandinMeshData> _d;
    int _f;
};

class Mesh {
public:
    Mesh() : _d(std::make_shared<StandinMeshData>()) {}
    explicit Mesh(std::shared_ptr<const StandinMeshData> d) : _d(std::move(d)) {}
    int getMeshDimension() const { return _d->dim; }
    int getNumberOfCells() const { return (int)_d->cell_faces.size(); }
    int getNumberOfFaces() const { return (int)_d->face_measure.size(); }
    Cell getCell(int j) const { return Cell(_d, j); }
    Face getFace(int f) const { return Face(_d, f); }
    int getIndexFacePeriodic(int f) const
    {
        const int t = _d->face_twin[(size_t)f];
        if (t < 0) throw CdmathException("getIndexFacePeriodic: the face has no periodic twin");
        return t;
    }
    double getXMin() const { return _d->lo[0]; }
    double getXMax() const { return _d->hi[0]; }
    double getYMin() const { return _d->lo[1]; }
    double getYMax() const { return _d->hi[1]; }
    double getZMin() const { return _d->lo[2]; }
    double getZMax() const { return _d->hi[2]; }

private:
    std::shared_ptr<const StandinMeshData> _d;
};

/* a cell field with one or several components; only what initial_conditions_shock touches */
class Field {
public:
    Field(int ncells, int ncomp = 1) : _nc(ncomp), _v((size_t)ncells * ncomp, 0.0) {}
    double &operator()(int j) { return _v[(size_t)j * _nc]; }
    double &operator()(int j, int c) { return _v[(size_t)j * _nc + c]; }
    double &operator[](int j) { return _v[(size_t)j]; }

private:
    int _nc;
    std::vector<double> _v;
};
#endif
