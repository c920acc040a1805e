package cuda

// The //export lives in its own file: cgo emits the preamble of a file that uses //export twice, so such a preamble may
// only contain declarations (cuda.go holds the C helper that takes this function's address).

/*
#include "ptb200.h"
*/
import "C"

import (
	"runtime/cgo"
	"unsafe"
)

//export ptbGoProgress
func ptbGoProgress(user unsafe.Pointer) {
	h := *(*cgo.Handle)(user)
	h.Value().(func())()
}
