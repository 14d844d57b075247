// Test-infrastructure stub: heikoburau/QuantumExpression is absent from this image and the
// arithmetic of the hot path does not live in it (it only builds operators); the shim feeds
// (coefficient, a-mask, b-mask) arrays directly. See oracle/ref_shim/shim.cu.
#pragma once
namespace quantum_expression { class PauliExpression; class FermionExpression; }
