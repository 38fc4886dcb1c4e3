// Lattice geometry: Bravais vectors, basis cell, number of cells, per-site mu_s and atom types.
// Mirrors the observable behaviour of Data::Geometry (core/include/data/Geometry.hpp:77-175,
// core/src/data/Geometry.cpp:22-160,486-800) without the triangulation (qhull, visualisation only).
//
// Site order is the reference's: ispin = ib + n_cell_atoms*(a + Na*(b + Nb*c)).
#pragma once

#include "types.hpp"

namespace sb
{

enum class BravaisLatticeType
{
    Irregular   = 0,
    Rectilinear = 1,
    SC          = 2,
    Hex2D       = 3,
    Hex2D60     = 4,
    Hex2D120    = 5,
    HCP         = 6,
    BCC         = 7,
    FCC         = 8
};

struct Geometry
{
    Geometry(
        const std::vector<Vec3> & bravais_vectors, const std::array<int, 3> & n_cells,
        const std::vector<Vec3> & cell_atoms, const std::vector<double> & cell_mu_s, double lattice_constant );

    static std::vector<Vec3> BravaisVectorsSC();
    static std::vector<Vec3> BravaisVectorsFCC();
    static std::vector<Vec3> BravaisVectorsBCC();
    static std::vector<Vec3> BravaisVectorsHex2D60();
    static std::vector<Vec3> BravaisVectorsHex2D120();

    std::vector<Vec3> bravais_vectors;
    double lattice_constant;
    std::array<int, 3> n_cells;
    int n_cell_atoms;
    std::vector<Vec3> cell_atoms;
    std::vector<double> cell_mu_s; // per basis atom
    std::vector<int> cell_atom_types;

    BravaisLatticeType classifier = BravaisLatticeType::Irregular;
    int nos;
    int nos_nonvacant;
    int n_cells_total;
    int dimensionality       = 0;
    int dimensionality_basis = 0;
    Vec3 center, bounds_min, bounds_max, cell_bounds_min, cell_bounds_max;

    // Position of basis atom `iatom` in cell (a,b,c); same expression as Geometry.cpp:150-154
    Vec3 position_of( std::int64_t a, std::int64_t b, std::int64_t c, int iatom ) const;

    // Per-site arrays, generated on first use (402 MB each at 256^3 -- not needed by the device path)
    const vectorfield & positions() const;
    const scalarfield & mu_s() const;
    const intfield & atom_types() const;

    bool mu_s_homogeneous() const;

private:
    void calculateBounds();
    void calculateUnitCellBounds();
    void calculateDimensionality();
    void calculateGeometryType();

    mutable vectorfield _positions;
    mutable scalarfield _mu_s;
    mutable intfield _atom_types;
};

} // namespace sb
