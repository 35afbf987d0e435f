// cusim stub
