// Host-side setup of interaction pair lists: neighbour shells and DMI normals.
// Behaviour follows Engine::Neighbours (core/src/engine/Neighbours.cpp:15-156,249-286) so that the
// generated pair lists are identical to the reference's (order included).
#pragma once

#include "geometry.hpp"

namespace sb
{
namespace neighbours
{

std::vector<double> get_shell_radii( const Geometry & geometry, std::size_t n_shells );

void get_neighbours_in_shells(
    const Geometry & geometry, std::size_t n_shells, pairfield & neighbours, intfield & shells,
    bool use_redundant_neighbours );

Vec3 dmi_normal_from_pair( const Geometry & geometry, const Pair & pair, int chirality );

} // namespace neighbours
} // namespace sb
