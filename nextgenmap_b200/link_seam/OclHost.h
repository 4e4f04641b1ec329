// link_seam/OclHost.h -- zero-patch link seam for NextGenMap.
//
// NGM's factory (_NGM::CreateAlignment / DeleteAlignment, reference src/NGM.cpp:388-437) is the only
// place that names the alignment backend: it constructs `OclHost(dev_type, id, threads)` and
// `SWOclCigar(host)` and later calls `((SWOcl*) instance)->getHost()`.  Putting this directory in
// front of lib/mason/opencl on NGM's include path makes NGM.cpp compile against the classes below,
// which forward to libngm_b200.so -- no NGM source file is modified.
#ifndef NGM_B200_LINK_SEAM_OCLHOST_H
#define NGM_B200_LINK_SEAM_OCLHOST_H

#ifndef CL_DEVICE_TYPE_CPU
#define CL_DEVICE_TYPE_CPU (1 << 1)
#define CL_DEVICE_TYPE_GPU (1 << 2)
#endif

class OclHost {
public:
	OclHost(int const device_type, int gpu_id, int const cpu_cores);
	virtual ~OclHost();
	int cudaDevice() const { return device; }

private:
	int device;
};

#endif
