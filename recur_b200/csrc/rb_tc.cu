/* rb_tc.cu placeholder: tensor-core engine lands here */
