// TEST INFRASTRUCTURE (oracle build shim) -- not product code.
// Shadows Storm/Bittern/MatrixDense.hpp for the mesh-less solver TU only:
// Storm/Solvers/MatrixDense.hpp:43-46 declares a legacy DenseMatrix that
// collides with Bittern's. Intentionally empty.
#pragma once
