// module.h -- base class of every flow module (operator).
// Interface of cudarecv/modules/inc/module.h:13-144: Start/Update/Stop with the flow's stream
// passed as void*, named input/output ports, typed named parameters.  0 = ok, non-zero = fatal.
#ifndef DPE_HOST_MODULE_H_
#define DPE_HOST_MODULE_H_

#include <map>
#include <string>
#include <vector>
#include "dsp.h"

namespace dsp {

class Module {
  public:
    virtual ~Module() {}
    virtual int Start(void* cuFlowStream) { (void)cuFlowStream; return 0; }
    virtual int Update(void* cuFlowStream) = 0;
    virtual int Stop() { return 0; }

    std::string GetModuleName() const { return ModuleName; }

    int GetInputID(const char* name) const;
    int GetOutputID(const char* name) const;
    int SetInput(unsigned char id, Port* in);          // type / length checked like module.cpp:284-310
    int GetOutput(unsigned char id, Port** out);
    int NumInputPorts() const { return (int)expectedInputs.size(); }
    int NumOutputPorts() const { return (int)outputs.size(); }
    const Port* OutputAt(int id) const { return (id >= 0 && id < (int)outputs.size()) ? &outputs[id] : nullptr; }

    int SetParam(const std::string& key, const int val);
    int SetParam(const std::string& key, const char val);
    int SetParam(const std::string& key, const float val);
    int SetParam(const std::string& key, const double val);
    int SetParam(const std::string& key, const bool val);
    int SetParam(const std::string& key, const char* str);
    int GetParam(const std::string& key, int* val) const;
    int GetParam(const std::string& key, float* val) const;
    int GetParam(const std::string& key, double* val) const;
    int GetParam(const std::string& key, bool* val) const;
    int GetParam(const std::string& key, char* str, unsigned int capacity) const;

  protected:
    std::string ModuleName;
    std::map<std::string, Param> Params;
    std::vector<ExpectedPort> expectedInputs;
    std::vector<Port*> inputs;
    std::vector<Port> outputs;

    int InsertParam(const std::string& key, void* ptr, DataType_t dtype, unsigned int capacity, unsigned int size);
    void AllocateInputs(unsigned char n) { expectedInputs.resize(n); inputs.assign(n, nullptr); }
    void AllocateOutputs(unsigned char n) { outputs.resize(n); }
    int ConfigExpectedInput(unsigned char id, const char* name, DataType_t dtype, ValueType_t vtype,
                            unsigned short vectorLength);
    int ConfigOutput(unsigned char id, const char* name, DataType_t dtype, ValueType_t vtype, MemLoc_t loc,
                     unsigned short vectorLength, void* data, int aux);
    int UpdateOutput(unsigned char id, int64_t length, void* data, int aux);
    bool InputsConnected() const;

    template <typename T> const T* In(int id) const { return static_cast<const T*>(inputs[id]->Data); }
    int64_t InLen(int id) const { return inputs[id]->Length; }

  private:
    int SetParamRaw(const std::string& key, DataType_t dtype, const void* src, unsigned int size);
    int GetParamRaw(const std::string& key, DataType_t dtype, void* dst, unsigned int size) const;
};

}  // namespace dsp
#endif
