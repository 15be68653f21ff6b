// gpu_train_{gcn,sage,gat}: the reference's CLI (src/gnn/train.cpp:8-42) over the B200 object set.
//   ./gpu_train_gcn cora 200 1 softmax [hidden score_drop feat_drop lr [layers subg_size val_interval inductive]]
#include <chrono>
#include <iostream>
#include "gai_model.h"

int main(int argc, char* argv[]) {
  if (argc <= 4 || (argc > 9 && argc != 13)) {
    std::cout << "Usage: ./train data num_epochs num_threads type_loss "
              << "hidden(16) score_drop_rate(0.) feat_drop_rate(0.) "
              << "learnng_rate(0.01) num_layers(2) subg_size(0) val_interval(50) inductive(0)\n"
              << "Example: ./gpu_train_gcn citeseer 10 2 softmax\n";
    return 1;
  }
#ifdef USE_GAT
  Model<GAT_layer> model;
  std::cout << "Using Graph Attention Network\n";
#elif defined(USE_SAGE)
  Model<SAGE_layer> model;
  std::cout << "Using GraphSAGE\n";
#else
  Model<GCN_layer> model;
  std::cout << "Using Graph Convolutional Network\n";
#endif
  model.load_data(argc, argv);
  model.construct_network();
  auto t1 = std::chrono::steady_clock::now();
  model.train();
  auto t2 = std::chrono::steady_clock::now();
  std::cout << "Total training time (validation time included): " << std::chrono::duration<double>(t2 - t1).count() << " seconds\n";
  auto test_acc = model.evaluate("test");
  auto t3 = std::chrono::steady_clock::now();
  std::cout << "Test accuracy: " << test_acc << "  test time: " << std::chrono::duration<double>(t3 - t2).count() << " seconds\n";
  return 0;
}
