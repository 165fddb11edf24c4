// TEST INFRASTRUCTURE (oracle build only) -- zero-initialising global allocator.
// The reference reads uninitialised heap on the hot path: `new marginal_information[n]`
// (algorithms/epistasis_func.cpp:713) leaves dPbc/dPca unset for zero-count genotype classes
// (genetics/genotype/common_genotype_func.cpp:202-215), and CompressedGenotypeTable5::selectCaseControl
// walks the mask arrays past their initialised part (compressed_genotype_table5.cpp:515). With a
// zeroing allocator the reference's output is deterministic; that deterministic behaviour is the
// parity target (SURVEY.md fact 5 / defect D1, D2).
#include <cstdlib>
#include <new>
void *operator new(std::size_t n) { void *p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void *operator new[](std::size_t n) { void *p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void *p) noexcept { std::free(p); }
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete(void *p, std::size_t) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }
