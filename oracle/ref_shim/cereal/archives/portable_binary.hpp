// stand-in for cereal (see ../../README.md): archives that do nothing
#pragma once
#include <istream>
#include <ostream>
namespace cereal {
struct PortableBinaryOutputArchive { PortableBinaryOutputArchive(std::ostream&) {} template <class... A> void operator()(A&&...) {} };
struct PortableBinaryInputArchive { PortableBinaryInputArchive(std::istream&) {} template <class... A> void operator()(A&&...) {} };
}  // namespace cereal
