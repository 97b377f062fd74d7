// Compatibility shim: consumers of lanxinger/spz include "load-spz.h"; the whole API lives in
// include/spz_b200/spz.hpp.  Add -I<repo>/include/spz to compile such a consumer unchanged.
#pragma once
#include "../spz_b200/spz.hpp"
