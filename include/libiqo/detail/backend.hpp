// Shared by the three public resizer headers of the CUDA backend.
// Declares the opaque C-ABI handle (include/iqo_cuda.h) the classes hold and the export macro.
#ifndef LIBIQO_DETAIL_BACKEND_HPP
#define LIBIQO_DETAIL_BACKEND_HPP

#include <stddef.h>

struct iqo_cuda_resizer;

#ifndef IQO_EXPORT
#  ifdef __GNUC__
#    define IQO_EXPORT __attribute__((visibility("default")))
#  else
#    define IQO_EXPORT
#  endif
#endif

#endif
