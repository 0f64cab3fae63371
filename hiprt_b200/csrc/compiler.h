// Runtime compilation of user trace kernels (NVRTC, sm_100a) and the few CUDA driver-API
// calls the HIPRT API needs (CUcontext binding, CUmodule/CUfunction handles). Both
// libraries are dlopen'ed on first use so the shared library loads on machines without a
// GPU driver (the CPU-side tests only check the exported symbols).
// Reference counterpart: hiprt/impl/Compiler.{h,cpp} (JIT + func-table codegen :573-625).
#pragma once
#include <hiprt/hiprt_types.h>

#include <stdexcept>

namespace hiprt
{
namespace compiler
{
struct NotAvailable : std::runtime_error
{
	using std::runtime_error::runtime_error;
};

void setCurrentContext( void* cuContext );

void buildTraceKernels(
	int				  ccMajor,
	int				  ccMinor,
	uint32_t		  numFunctions,
	const char**	  funcNames,
	const char*		  src,
	const char*		  moduleName,
	uint32_t		  numHeaders,
	const char**	  headers,
	const char**	  includeNames,
	uint32_t		  numOptions,
	const char**	  options,
	uint32_t		  numGeomTypes,
	uint32_t		  numRayTypes,
	hiprtFuncNameSet* funcNameSets,
	hiprtApiFunction* functionsOut,
	hiprtApiModule*	  moduleOut );

// embedded copy of include/hiprt/*.h needed by device code (generated at build time)
uint32_t deviceHeaders( const char*** namesOut, const char*** sourcesOut );
} // namespace compiler
} // namespace hiprt
