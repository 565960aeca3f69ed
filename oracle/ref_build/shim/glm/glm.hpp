// TEST INFRASTRUCTURE (oracle build shim) -- not product code.
// Storm/Feathers/Field.hpp:37,45-48 only aliases four glm types; nothing on the
// Krylov path touches them.
#pragma once
namespace glm {
struct dvec2 { double x, y; };
struct dvec3 { double x, y, z; };
struct dmat2 { double m[4]; };
struct dmat3 { double m[9]; };
} // namespace glm
