// Integration shim: shadows Storm/Bittern/MatrixDense.hpp in translation units that include the
// legacy solver headers. Storm/Solvers/MatrixDense.hpp:43-46 declares its own DenseMatrix /
// DenseVector, which collide with Bittern's (the solver headers do not compile otherwise,
// SURVEY.md F4/F5). Put this directory FIRST on the include path. Intentionally empty.
#pragma once
