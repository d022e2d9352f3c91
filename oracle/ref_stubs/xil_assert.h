#ifndef XIL_ASSERT_H
#define XIL_ASSERT_H
#include <stdlib.h>
#include <string.h>
#define Xil_AssertVoid(e) do { if (!(e)) abort(); } while (0)
#define Xil_AssertNonvoid(e) do { if (!(e)) abort(); } while (0)
#endif
