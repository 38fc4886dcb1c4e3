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

// A lattice site named by basis atom and cell translations (Data::Site, core/include/engine/Vectormath.hpp)
struct LatticeSite
{
    int i = 0;
    int translations[3] = { 0, 0, 0 };
};
// Pinned boundary cells and individually pinned sites (Data::Pinning, core/include/data/Geometry.hpp:40-58)
struct Pinning
{
    int na_left = 0, na_right = 0, nb_left = 0, nb_right = 0, nc_left = 0, nc_right = 0;
    std::vector<Vec3> pinned_cell;   // orientation of the pinned boundary cells, per basis atom
    std::vector<LatticeSite> sites;  // additional pinned sites ...
    std::vector<Vec3> spins;         // ... and their orientations
};
// Defect sites (Data::Defects, core/include/data/Geometry.hpp:60-66): type < 0 is a vacancy
struct Defects
{
    std::vector<LatticeSite> sites;
    std::vector<int> types;
};
// per-site flags of a lattice with pinned sites or defects
constexpr unsigned char SITE_VACANT  = 1; // atom type < 0: the site takes part in no interaction (check_atom_type, Vectormath.hpp:406-433)
constexpr unsigned char SITE_PINNED  = 2; // mask_unpinned == 0: force and virtual force are zero (Method_LLG.cpp:122-124, 222-224)
constexpr unsigned char SITE_NO_MU_S = 4; // mu_s == 0 (every defect site, Geometry.cpp:72-82): no Zeeman term, no dipolar moment

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

    // Pinning and defects (the reference's compile-time options SPIRIT_ENABLE_PINNING / SPIRIT_ENABLE_DEFECTS, always built
    // here). `site_flags` is empty for a lattice without either; otherwise one byte per site in the reference's site order.
    // `mask_pinned_cells` holds the orientation pinned sites are reset to (Geometry::Apply_Pinning, Geometry.cpp:807-832).
    // `site_revision` changes whenever the flags do (the device copy follows it).
    void set_pinning_and_defects( const Pinning & pinning, const Defects & defects );
    void set_pinned( int ispin, bool pinned, const Vec3 & orientation ); // Configurations::Set_Pinned, Configurations.cpp:583-599
    void set_atom_type( int ispin, int type );                           // Configurations::Set_Atom_Types, :567-581
    void set_vacancy_read_from_file( int ispin ); // a (near-)zero vector in a spin file: atom type -1, moment kept (IO.cpp:265-275)
    void apply_pinning( Vec3 * spins ) const;
    bool has_site_flags() const
    {
        return !site_flags.empty();
    }
    int site_index( const LatticeSite & site ) const;
    Pinning pinning;
    Defects defects;
    std::vector<unsigned char> site_flags;
    vectorfield mask_pinned_cells;
    std::uint64_t site_revision = 0;

private:
    void calculateBounds();
    void calculateUnitCellBounds();
    void calculateDimensionality();
    void calculateGeometryType();
    void need_site_flags();

    mutable vectorfield _positions;
    mutable scalarfield _mu_s;
    mutable intfield _atom_types;
};

} // namespace sb
