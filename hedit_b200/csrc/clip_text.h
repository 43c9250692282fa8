// CLIP text tower (forward only); see clip_text.cu.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "netexec.h"

namespace hedit {

struct TextCfg { int vocab = 49408, width = 768, heads = 12, layers = 12, ffn = 3072, tokens = 77; };

class ClipText : public NetExec {
 public:
  explicit ClipText(const TextCfg& cfg);
  ~ClipText();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  // returns 1 for tensors this path does not use (position_ids, ...)
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st);
  int finalize(std::string* missing);
  // ids [B][tokens] int32 (host) -> last_hidden_state [B][tokens][width] fp32 (device)
  int forward(const int32_t* ids_host, int B, float* out, cudaStream_t st);

 private:
  struct Slot { std::vector<int64_t> shape; int kind = 0; void* dst = nullptr; int ld = 0, off = 0; bool loaded = false; };   // 0 f32, 2 rows
  struct Layer { float *ln1g = 0, *ln1b = 0, *ln2g = 0, *ln2b = 0, *b_qkv = 0, *b_o = 0, *b_fc1 = 0, *b_fc2 = 0; op_t *w_qkv = 0, *w_o = 0, *w_fc1 = 0, *w_fc2 = 0; };
  template <typename T> T* walloc(size_t n);
  void reg(const std::string& name, std::vector<int64_t> shape, int kind, void* dst, int ld, int off);
  int run(const int* ids_dev, int B, float* out);

  TextCfg cfg_;
  std::map<std::string, Slot> slots_;
  std::vector<void*> owned_;
  float *tok_ = 0, *pos_ = 0, *lnf_g_ = 0, *lnf_b_ = 0, *stage_ = 0;
  std::vector<Layer> layers_;
  int* ids_dev_ = nullptr; int ids_cap_ = 0;
};

}  // namespace hedit
