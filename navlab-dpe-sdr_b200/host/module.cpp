#include "module.h"
#include <cstring>
#include <iostream>

namespace dsp {

static unsigned short clamp16(int64_t n) { return n > 65535 ? 65535 : (unsigned short)n; }

int Module::GetInputID(const char* name) const {
    for (size_t i = 0; i < expectedInputs.size(); ++i)
        if (std::strcmp(expectedInputs[i].Name, name) == 0) return (int)i;
    std::cerr << "[" << ModuleName << "] no input port named " << name << std::endl;
    return -1;
}

int Module::GetOutputID(const char* name) const {
    for (size_t i = 0; i < outputs.size(); ++i)
        if (std::strcmp(outputs[i].Name, name) == 0) return (int)i;
    std::cerr << "[" << ModuleName << "] no output port named " << name << std::endl;
    return -1;
}

int Module::SetInput(unsigned char id, Port* in) {
    if (id >= expectedInputs.size() || !in) return -1;
    const ExpectedPort& e = expectedInputs[id];
    if (e.ValueType != VALUETYPE_ANY && e.ValueType != in->ValueType) {
        std::cerr << "[" << ModuleName << "] input " << e.Name << ": value type mismatch with " << in->Name << std::endl;
        return -1;
    }
    if (e.VectorLength != VECTORLENGTH_ANY && in->VectorLength != VECTORLENGTH_ANY &&
        e.VectorLength != in->VectorLength) {
        std::cerr << "[" << ModuleName << "] input " << e.Name << ": vector length mismatch with " << in->Name << std::endl;
        return -1;
    }
    inputs[id] = in;
    return 0;
}

int Module::GetOutput(unsigned char id, Port** out) {
    if (id >= outputs.size() || !out) return -1;
    *out = &outputs[id];
    return 0;
}

int Module::InsertParam(const std::string& key, void* ptr, DataType_t dtype, unsigned int capacity, unsigned int size) {
    Param p = {ptr, dtype, capacity, size};
    Params[key] = p;
    return 0;
}

int Module::ConfigExpectedInput(unsigned char id, const char* name, DataType_t dtype, ValueType_t vtype,
                                unsigned short vectorLength) {
    if (id >= expectedInputs.size()) return -1;
    ExpectedPort& e = expectedInputs[id];
    std::strncpy(e.Name, name, sizeof(e.Name) - 1);
    e.Name[sizeof(e.Name) - 1] = 0;
    e.Datatype = dtype; e.ValueType = vtype; e.VectorLength = vectorLength;
    return 0;
}

int Module::ConfigOutput(unsigned char id, const char* name, DataType_t dtype, ValueType_t vtype, MemLoc_t loc,
                         unsigned short vectorLength, void* data, int aux) {
    if (id >= outputs.size()) return -1;
    Port& p = outputs[id];
    std::memset(&p, 0, sizeof(p));
    std::strncpy(p.Name, name, sizeof(p.Name) - 1);
    p.Datatype = dtype; p.ValueType = vtype; p.MemLoc = loc; p.VectorLength = vectorLength;
    p.Length = vectorLength; p.Data = data; p.AuxValue = aux;
    return 0;
}

int Module::UpdateOutput(unsigned char id, int64_t length, void* data, int aux) {
    if (id >= outputs.size()) return -1;
    outputs[id].Length = length;
    outputs[id].VectorLength = clamp16(length);
    outputs[id].Data = data;
    outputs[id].AuxValue = aux;
    return 0;
}

bool Module::InputsConnected() const {
    for (size_t i = 0; i < inputs.size(); ++i)
        if (!inputs[i]) {
            std::cerr << "[" << ModuleName << "] input " << expectedInputs[i].Name << " is not connected" << std::endl;
            return false;
        }
    return true;
}

// dtype + capacity check, then copy (module.cpp:21-114 of the reference)
int Module::SetParamRaw(const std::string& key, DataType_t dtype, const void* src, unsigned int size) {
    std::map<std::string, Param>::iterator it = Params.find(key);
    if (it == Params.end()) {
        std::cerr << "[" << ModuleName << "] SetParam: no parameter " << key << std::endl;
        return -1;
    }
    Param& p = it->second;
    if (p.Datatype != dtype) {
        std::cerr << "[" << ModuleName << "] SetParam(" << key << "): datatype mismatch" << std::endl;
        return -1;
    }
    if (size > p.Capacity) {
        std::cerr << "[" << ModuleName << "] SetParam(" << key << "): value does not fit" << std::endl;
        return -1;
    }
    std::memcpy(p.Ptr, src, size);
    p.Size = size;
    return 0;
}

int Module::GetParamRaw(const std::string& key, DataType_t dtype, void* dst, unsigned int size) const {
    std::map<std::string, Param>::const_iterator it = Params.find(key);
    if (it == Params.end() || it->second.Datatype != dtype || size > it->second.Capacity) return -1;
    std::memcpy(dst, it->second.Ptr, size);
    return 0;
}

int Module::SetParam(const std::string& k, const int v) { return SetParamRaw(k, INT_t, &v, sizeof(v)); }
int Module::SetParam(const std::string& k, const char v) { return SetParamRaw(k, CHAR_t, &v, sizeof(v)); }
int Module::SetParam(const std::string& k, const float v) { return SetParamRaw(k, FLOAT_t, &v, sizeof(v)); }
int Module::SetParam(const std::string& k, const double v) { return SetParamRaw(k, DOUBLE_t, &v, sizeof(v)); }
int Module::SetParam(const std::string& k, const bool v) { return SetParamRaw(k, BOOL_t, &v, sizeof(v)); }
int Module::SetParam(const std::string& k, const char* s) {
    return SetParamRaw(k, CHAR_t, s, (unsigned int)std::strlen(s) + 1);
}
int Module::GetParam(const std::string& k, int* v) const { return GetParamRaw(k, INT_t, v, sizeof(*v)); }
int Module::GetParam(const std::string& k, float* v) const { return GetParamRaw(k, FLOAT_t, v, sizeof(*v)); }
int Module::GetParam(const std::string& k, double* v) const { return GetParamRaw(k, DOUBLE_t, v, sizeof(*v)); }
int Module::GetParam(const std::string& k, bool* v) const { return GetParamRaw(k, BOOL_t, v, sizeof(*v)); }
int Module::GetParam(const std::string& k, char* s, unsigned int capacity) const {
    std::map<std::string, Param>::const_iterator it = Params.find(k);
    if (it == Params.end() || it->second.Datatype != CHAR_t || it->second.Size > capacity) return -1;
    std::memcpy(s, it->second.Ptr, it->second.Size);
    return 0;
}

}  // namespace dsp
