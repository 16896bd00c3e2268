/* Stand-in for SDAR's Common/Float.h (an un-vendored submodule of the reference): the reference's
 * src/changeover.hpp only needs the Float typedef from it.  Used when compiling oracle/_ref. */
#pragma once
typedef double Float;
