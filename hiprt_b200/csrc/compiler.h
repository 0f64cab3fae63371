// Runtime compilation of user trace kernels (NVRTC, sm_100a) and the few CUDA driver-API
// calls the HIPRT API needs (CUcontext binding, CUmodule/CUfunction handles). Both
// libraries are dlopen'ed on first use so the shared library loads on machines without a
// GPU driver (the CPU-side tests only check the exported symbols).
// Reference counterpart: hiprt/impl/Compiler.{h,cpp} (JIT + func-table codegen :573-625).
#pragma once
#include <hiprt/hiprt_types.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace hiprt
{
namespace compiler
{
struct NotAvailable : std::runtime_error
{
	using std::runtime_error::runtime_error;
};

void setCurrentContext( void* cuContext );

// Loaded modules of ONE hiprtContext, keyed by source + options (+ function table). CUmodule / CUfunction handles belong to the
// CUDA context they were loaded in, so the cache lives in the hiprtContext (as the reference's does: Compiler is a member of
// Context) and is unloaded with it.
class ModuleCache
{
  public:
	ModuleCache() = default;
	ModuleCache( const ModuleCache& ) = delete;
	ModuleCache& operator=( const ModuleCache& ) = delete;
	~ModuleCache();
	void*  find( const std::string& key );
	void*  insert( const std::string& key, void* module ); // returns the module kept (an earlier one if another thread won)
	size_t size();

  private:
	std::map<std::string, void*> m_modules;
	std::mutex					 m_mutex;
};

void buildTraceKernels(
	ModuleCache&	  cache,
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

// Offline-compiled user kernels (reference: Compiler::buildKernelsFromBitcode, Compiler.cpp:338-470): `bitcode` is the caller's
// relocatable device code for this architecture (nvcc -dc: cubin, PTX, fatbin or LTO-IR) built against <hiprt/hiprt_device.h>
// with intersectFunc / filterFunc left undefined; they are generated here from the function-name table, compiled as relocatable
// code and linked in with nvJitLink. linkTraceKernels is the device-independent half (needs no GPU): the linked cubin.
std::vector<char> linkTraceKernels(
	int ccMajor, int ccMinor, const char* bitcode, size_t bitcodeSize, const char* moduleName, uint32_t numGeomTypes, uint32_t numRayTypes, hiprtFuncNameSet* funcNameSets );

void buildTraceKernelsFromBitcode(
	ModuleCache&	  cache,
	int				  ccMajor,
	int				  ccMinor,
	uint32_t		  numFunctions,
	const char**	  funcNames,
	const char*		  moduleName,
	const char*		  bitcode,
	size_t			  bitcodeSize,
	uint32_t		  numGeomTypes,
	uint32_t		  numRayTypes,
	hiprtFuncNameSet* funcNameSets,
	hiprtApiFunction* functionsOut );

// embedded copy of include/hiprt/*.h needed by device code (generated at build time)
uint32_t deviceHeaders( const char*** namesOut, const char*** sourcesOut );
} // namespace compiler
} // namespace hiprt
