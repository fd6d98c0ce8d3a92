// dpeflow.h -- the DPE flow: DPInit -> SampleBlock -> BatchCorrScores -> BatchCorrManifold ->
// cuEKF -> cuChanMgr -> XECEFLogger (cudarecv/dsp/inc/dpeflow.h, dsp/src/dpeflow.cpp:26-222).
#ifndef DPE_HOST_DPEFLOW_H_
#define DPE_HOST_DPEFLOW_H_
#include "flow.h"

namespace dsp {
class DPEFlow : public Flow {
  public:
    ~DPEFlow() override {}
    /** Builds the module graph with the reference's default parameters; `filename` is unused
     *  (as in the reference).  File paths default to $HOME/Desktop/demofiles/... and are meant to
     *  be overridden with setparam / SetModParam before startflow. */
    int LoadFlow(const char* filename) override;
};
}  // namespace dsp
#endif
